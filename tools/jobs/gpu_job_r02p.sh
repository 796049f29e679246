#!/bin/bash
# Round-2 GPU job "p": tests, A/B of the scaled-domain element path (FP8FQ_MAGIC), channel-innermost shapes, MSE, bench, ncu.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"
timeout 900 python tools/ab_build_options.py --only nomagic > gpurun_out/r02p_ab.log 2>&1; echo "ab rc=$?"
cp gpurun_out/ab_build_options.json gpurun_out/ab_build_options_r02p.json
for mv in 3.0; do
  CL_MAXVAL=$mv CL_JSON=cl_shapes_r02p_mv3_default.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02p_cl_mv3_default.log 2>&1
  FP8FQ_LIB=$PWD/build_variants/libfp8fq_nomagic.so CL_MAXVAL=$mv CL_JSON=cl_shapes_r02p_mv3_nomagic.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02p_cl_mv3_nomagic.log 2>&1
done
MSE_JSON=mse_r02p_default.json timeout 300 python tools/bench_mse.py > /dev/null 2>&1
FP8FQ_LIB=$PWD/build_variants/libfp8fq_nomagic.so MSE_JSON=mse_r02p_nomagic.json timeout 300 python tools/bench_mse.py > /dev/null 2>&1
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02p.json 2> gpurun_out/bench_r02p.err; echo "bench rc=$?"
CL_MAXVAL=3.0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fq_ -o gpurun_out/prof_mbv2_r02p -f python tools/profile_targets_mbv2.py > gpurun_out/ncu_mbv2_r02p.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r02p_pytest.log
