#!/bin/bash
# Round-2 final verification: GPU tests, smoke(), bench (both layouts + reference arm), launch list, ncu full summaries, site shapes.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02final_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02final_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02final.json 2> gpurun_out/bench_r02final.err; echo "bench rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02final.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-model --no-configs > gpurun_out/launches_r02final.log 2>&1; echo "launch list rc=$?"
CL_MAXVAL=3.0 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:fq_ -o gpurun_out/prof_mbv2_r02final -f python tools/profile_targets_mbv2.py > gpurun_out/ncu_mbv2_r02final.log 2>&1
python tools/summarize_ncu.py full gpurun_out/prof_mbv2_r02final.ncu-rep gpurun_out/ncu_full_mbv2_r02final.json > /dev/null 2>&1 && rm -f gpurun_out/prof_mbv2_r02final.ncu-rep
tail -2 gpurun_out/r02final_pytest.log; tail -1 gpurun_out/r02final_smoke.log
