#!/bin/bash
# Round-2 GPU job "v": scaled-domain constants at fixed table offsets (independent prologue loads); default vs two0 vs look-up
# build on config 3 and the channel-innermost site shapes; GPU tests.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"
for v in default two0 nomagic; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  C3_JSON=c3_r02v_$v.json timeout 300 python tools/bench_c3.py > gpurun_out/r02v_c3_$v.log 2>&1; echo "c3 $v rc=$?"
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02v_mv3_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02v_cl_mv3_$v.log 2>&1
  CL_MAXVAL=4.0 CL_JSON=cl_shapes_r02v_mv4_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02v_cl_mv4_$v.log 2>&1
done
unset FP8FQ_LIB
tail -2 gpurun_out/r02v_pytest.log
