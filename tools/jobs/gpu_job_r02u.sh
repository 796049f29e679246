#!/bin/bash
# Round-2 GPU job "u": two-group tables on the scaled-domain path (default) vs on the look-up path (two0) vs one loop (onepath),
# judged on config 3 itself (MobileNetV2 step + forward) and on the channel-innermost site shapes.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
for v in default two0 onepath; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  C3_JSON=c3_r02u_$v.json timeout 300 python tools/bench_c3.py > gpurun_out/r02u_c3_$v.log 2>&1; echo "c3 $v rc=$?"
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02u_mv3_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02u_cl_mv3_$v.log 2>&1
  CL_MAXVAL=4.0 CL_JSON=cl_shapes_r02u_mv4_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02u_cl_mv4_$v.log 2>&1
done
unset FP8FQ_LIB
echo done
