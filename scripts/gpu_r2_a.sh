#!/bin/bash
# round-2 GPU call A: k_team first contact + sanity of the round-1 head on this box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a/smi.txt 2>&1
python scripts/r2_team_sweep.py r2a > gpurun_out/r2a/sweep.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-sweep > gpurun_out/r2a/bench_head.json 2> gpurun_out/r2a/bench_head.err
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest_gpu.log 2>&1
tail -5 gpurun_out/r2a/pytest_gpu.log
tail -30 gpurun_out/r2a/sweep.log
