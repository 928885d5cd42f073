#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2t
timeout 200 ubench/red_probe 28 > gpurun_out/r2t/red_probe.jsonl 2>&1; tail -3 gpurun_out/r2t/red_probe.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > gpurun_out/r2t/pytest_gpu.log 2>&1
tail -4 gpurun_out/r2t/pytest_gpu.log | cut -c1-300
AGC_STAGE_THREADS=8 timeout 300 python scripts/r2_e2e.py 28 2>> gpurun_out/r2t/e2e.err | tee -a gpurun_out/r2t/e2e.jsonl
