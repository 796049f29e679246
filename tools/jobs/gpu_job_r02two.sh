#!/bin/bash
# two-group tables on the scaled-domain path with the group test on a flag bit (uniform branch): two1 vs default
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
for v in default two1; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  C3_JSON=c3_r02two_$v.json timeout 300 python tools/bench_c3.py > gpurun_out/r02two_c3_$v.log 2>&1; echo "c3 $v rc=$?"
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02two_mv3_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02two_cl_mv3_$v.log 2>&1
  CL_MAXVAL=4.0 CL_JSON=cl_shapes_r02two_mv4_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02two_cl_mv4_$v.log 2>&1
  python tools/ab_build_options.py --hash-leg > gpurun_out/r02two_hash_$v.log 2>&1
done
unset FP8FQ_LIB
echo done
