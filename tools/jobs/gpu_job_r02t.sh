#!/bin/bash
# Round-2 GPU job "t": cold paths out of line (FP8FQ_COLD_CALL) vs inlined.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"
timeout 1500 python tools/ab_build_options.py --only nomagic,nocold > gpurun_out/r02t_ab.log 2>&1; echo "ab rc=$?"
cp gpurun_out/ab_build_options.json gpurun_out/ab_build_options_r02t.json
for v in default nomagic nocold; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02t_mv3_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02t_cl_mv3_$v.log 2>&1
done
unset FP8FQ_LIB
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02t.json 2> gpurun_out/bench_r02t.err; echo "bench rc=$?"
tail -3 gpurun_out/r02t_pytest.log
