#!/bin/bash
# tiles per CTA for the K > 3 batch-norm kernels (default 1): FP8FQ_TILES_PER_CTA = 2 / 4 on the site shapes and on config 3
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
for t in 0 2 4; do
  if [ $t = 0 ]; then unset FP8FQ_TILES_PER_CTA; else export FP8FQ_TILES_PER_CTA=$t; fi
  CL_MAXVAL=3.0 CL_JSON=cl_shapes_tpc_$t.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/tpc_cl_$t.log 2>&1
  C3_JSON=c3_tpc_$t.json timeout 300 python tools/bench_c3.py > gpurun_out/tpc_c3_$t.log 2>&1
done
echo done
