#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2j/pytest_gpu.log 2>&1
tail -4 gpurun_out/r2j/pytest_gpu.log | cut -c1-220
timeout 300 python scripts/r2_scan_bench.py 28 > gpurun_out/r2j/scan.jsonl 2> gpurun_out/r2j/scan.err; cat gpurun_out/r2j/scan.jsonl; tail -2 gpurun_out/r2j/scan.err
AGC_SCAN_LEAN=0 timeout 300 python scripts/r2_scan_bench.py 28 2>/dev/null | grep '"chain": 1' > gpurun_out/r2j/scan_nolean.jsonl; cat gpurun_out/r2j/scan_nolean.jsonl
timeout 600 python bench.py --config C4 --steps 5 --warmup 3 > gpurun_out/r2j/bench_C4.json 2> gpurun_out/r2j/bench_C4.err; tail -c 1200 gpurun_out/r2j/bench_C4.json
timeout 200 python scripts/one_scan.py 26 > gpurun_out/r2j/plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_seg_scan_lean -s 2 -c 1 -o gpurun_out/r2j/prof_scan_lean python scripts/one_scan.py 26 > gpurun_out/r2j/ncu_scan.log 2>&1
tail -2 gpurun_out/r2j/ncu_scan.log
