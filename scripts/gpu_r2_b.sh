#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2b
python scripts/r2_team_sweep_b.py r2b > gpurun_out/r2b/sweep.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.log 2>&1
tail -15 gpurun_out/r2b/pytest_gpu.log
timeout 300 python scripts/r2_scan_bench.py 28 > gpurun_out/r2b/scan.jsonl 2> gpurun_out/r2b/scan.err
cat gpurun_out/r2b/scan.jsonl
