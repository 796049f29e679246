#!/bin/bash
# Round-2 GPU job "y": last verification of the final build (row-kernel entry points split): GPU tests, smoke(), bench.
cd "$(dirname "$0")/../.." || exit 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02y_pytest.log 2>&1; echo "pytest rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02y_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r02y.json 2> gpurun_out/bench_r02y.err; echo "bench rc=$?"
tail -2 gpurun_out/r02y_pytest.log; tail -1 gpurun_out/r02y_smoke.log
