#!/bin/bash
# Round-2 GPU job "w": verification of the final build -- GPU tests, smoke(), bench, launch list of the timed steps, ncu full.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02w_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02w.json 2> gpurun_out/bench_r02w.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 --memory-format channels_last --no-cpu --no-configs > gpurun_out/bench_r02w_channels_last.json 2> /dev/null; echo "bench cl rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02w_reference_arm.json 2> /dev/null; echo "ref arm rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02w.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-model --no-configs > gpurun_out/launches_r02w.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:fq_ -o gpurun_out/prof_targets_r02w -f python tools/profile_targets.py > gpurun_out/ncu_targets_r02w.log 2>&1
python tools/summarize_ncu.py full gpurun_out/prof_targets_r02w.ncu-rep gpurun_out/ncu_full_targets_r02w.json > /dev/null 2>&1 && rm -f gpurun_out/prof_targets_r02w.ncu-rep
CL_MAXVAL=3.0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fq_ -o gpurun_out/prof_mbv2_r02w -f python tools/profile_targets_mbv2.py > gpurun_out/ncu_mbv2_r02w.log 2>&1
python tools/summarize_ncu.py full gpurun_out/prof_mbv2_r02w.ncu-rep gpurun_out/ncu_full_mbv2_r02w.json > /dev/null 2>&1
CL_MAXVAL=3.0 CL_JSON=cl_shapes_r02w_mv3.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02w_cl_mv3.log 2>&1
CL_MAXVAL=4.0 CL_JSON=cl_shapes_r02w_mv4.json timeout 300 python tools/bench_cl_shapes.py > gpurun_out/r02w_cl_mv4.log 2>&1
MSE_JSON=mse_r02w.json timeout 300 python tools/bench_mse.py > /dev/null 2>&1
KERNELS_JSON=kernels_r02w.json timeout 300 python tools/bench_kernels.py 128 > /dev/null 2>&1
tail -2 gpurun_out/r02w_pytest.log; tail -2 gpurun_out/r02w_smoke.log
