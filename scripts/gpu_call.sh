#!/usr/bin/env bash
# Round-end check on a GPU box:  gpurun --timeout 1800 -- 'bash scripts/gpu_call.sh'
# (full GPU test-suite, the smoke entry point, the default headline bench; outputs under gpurun_out/)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:warnings > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-1700
