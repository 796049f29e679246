#!/bin/bash
# Round-2 GPU job "r": A/B of the options of the scaled-domain path (one loop instantiation, self-contained tie-guard lanes).
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
timeout 1500 python tools/ab_build_options.py --only nomagic,onepath,selfslow,onepath_selfslow > gpurun_out/r02r_ab.log 2>&1; echo "ab rc=$?"
cp gpurun_out/ab_build_options.json gpurun_out/ab_build_options_r02r.json
for v in default nomagic onepath selfslow onepath_selfslow; do
  if [ $v = default ]; then unset FP8FQ_LIB; else export FP8FQ_LIB=$PWD/build_variants/libfp8fq_$v.so; fi
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02r_mv3_$v.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02r_cl_mv3_$v.log 2>&1
done
unset FP8FQ_LIB
echo done
