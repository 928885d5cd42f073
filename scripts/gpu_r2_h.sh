#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2h
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2h/pytest_gpu.log 2>&1
tail -8 gpurun_out/r2h/pytest_gpu.log | cut -c1-220
timeout 600 python bench.py --config C5 --steps 5 --warmup 3 > gpurun_out/r2h/bench_C5.json 2> gpurun_out/r2h/bench_C5.err; tail -c 900 gpurun_out/r2h/bench_C5.json
timeout 600 python scripts/r2_multi_bench.py > gpurun_out/r2h/multi.jsonl 2> gpurun_out/r2h/multi.err; cat gpurun_out/r2h/multi.jsonl; tail -2 gpurun_out/r2h/multi.err
