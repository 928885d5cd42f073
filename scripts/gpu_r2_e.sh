#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2e/pytest_gpu.log 2>&1
tail -45 gpurun_out/r2e/pytest_gpu.log | cut -c1-220
timeout 300 python scripts/r2_scan_bench.py 28 > gpurun_out/r2e/scan.jsonl 2> gpurun_out/r2e/scan.err
cat gpurun_out/r2e/scan.jsonl; tail -3 gpurun_out/r2e/scan.err
for cfg in C3 C4 C5; do timeout 600 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r2e/bench_$cfg.json 2> gpurun_out/r2e/bench_$cfg.err; tail -c 2500 gpurun_out/r2e/bench_$cfg.json; tail -2 gpurun_out/r2e/bench_$cfg.err; done
timeout 200 python scripts/one_scan.py 26 > gpurun_out/r2e/plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_seg_scan_direct -s 2 -c 1 -o gpurun_out/r2e/prof_scan python scripts/one_scan.py 26 > gpurun_out/r2e/ncu_scan.log 2>&1
tail -3 gpurun_out/r2e/ncu_scan.log
