#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2g
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2g/pytest_gpu.log 2>&1
tail -8 gpurun_out/r2g/pytest_gpu.log | cut -c1-220
for sl in 0 64 256 1024; do AGC_LB_SLEEP=$sl timeout 300 python scripts/r2_scan_bench.py 28 2> gpurun_out/r2g/scan.err | grep '"chain": 1' | sed "s/^{/{\"sleep\": $sl, /" >> gpurun_out/r2g/scan_sleep.jsonl; done
cat gpurun_out/r2g/scan_sleep.jsonl
timeout 600 python bench.py --config C5 --steps 5 --warmup 3 > gpurun_out/r2g/bench_C5.json 2> gpurun_out/r2g/bench_C5.err; tail -c 1200 gpurun_out/r2g/bench_C5.json
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2g/smoke.log 2>&1; tail -2 gpurun_out/r2g/smoke.log
