#!/bin/bash
# Round-2 GPU job "s": element path decided once per launch (FP8FQ_MAGIC_HOIST) vs per vector; row-kernel registers.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"
timeout 1500 python tools/ab_build_options.py --only nomagic,nohoist,rows_minb1 > gpurun_out/r02s_ab.log 2>&1; echo "ab rc=$?"
cp gpurun_out/ab_build_options.json gpurun_out/ab_build_options_r02s.json
for v in default nomagic nohoist; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02s_mv3_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02s_cl_mv3_$v.log 2>&1
done
unset FP8FQ_LIB
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02s.json 2> gpurun_out/bench_r02s.err; echo "bench rc=$?"
tail -3 gpurun_out/r02s_pytest.log
