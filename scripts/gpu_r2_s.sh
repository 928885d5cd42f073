#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2s
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2s/pytest_gpu.log 2>&1
tail -4 gpurun_out/r2s/pytest_gpu.log | cut -c1-300
for cfg in "4 0" "4 1" "8 0" "8 1" "12 1" "16 1"; do
  set -- $cfg
  AGC_STAGE_THREADS=$1 AGC_NARROW_LABELS=$2 timeout 300 python scripts/r2_e2e.py 28 2>> gpurun_out/r2s/e2e.err | tee -a gpurun_out/r2s/e2e.jsonl
done
