#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2c
T=ubench/team_probe
for dbg in 0 1 2 3; do timeout 60 $T 2 0 0 30 60000 199552 0 3 0 0 $dbg >> gpurun_out/r2c/diag.jsonl 2>&1; done
for dbg in 0 1 2 3; do timeout 60 $T 4 0 0 30 100000 199552 0 3 0 0 $dbg >> gpurun_out/r2c/diag.jsonl 2>&1; done
for mc in 37 18; do timeout 60 $T 2 0 0 30 60000 199552 0 3 $mc 0 0 >> gpurun_out/r2c/diag.jsonl 2>&1; done
cat gpurun_out/r2c/diag.jsonl
timeout 120 $T 2 0 0 28 60000 199552 0 1 > gpurun_out/r2c/plain_team.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_team -c 1 -o gpurun_out/r2c/prof_team $T 2 0 0 28 60000 199552 0 1 > gpurun_out/r2c/ncu_team.log 2>&1
tail -3 gpurun_out/r2c/ncu_team.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest_gpu.log 2>&1
tail -15 gpurun_out/r2c/pytest_gpu.log
