#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2f
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2f/pytest_gpu.log 2>&1
tail -25 gpurun_out/r2f/pytest_gpu.log | cut -c1-220
timeout 300 python scripts/r2_scan_bench.py 28 > gpurun_out/r2f/scan.jsonl 2> gpurun_out/r2f/scan.err
cat gpurun_out/r2f/scan.jsonl; tail -3 gpurun_out/r2f/scan.err
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err
tail -c 1500 gpurun_out/r2f/bench.json; tail -3 gpurun_out/r2f/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2f/bench_reference.json 2> gpurun_out/r2f/bench_reference.err
tail -c 800 gpurun_out/r2f/bench_reference.json
timeout 300 python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > gpurun_out/r2f/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > gpurun_out/r2f/ncu_launches.log 2>&1
tail -2 gpurun_out/r2f/ncu_launches.log
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -o gpurun_out/r2f/prof_fast python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > gpurun_out/r2f/ncu_fast.log 2>&1
tail -2 gpurun_out/r2f/ncu_fast.log
