#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2x
timeout 600 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted" > gpurun_out/r2x/pytest_scan.log 2>&1
tail -3 gpurun_out/r2x/pytest_scan.log | cut -c1-300
timeout 300 python scripts/r2_scan_bench.py 28 2> gpurun_out/r2x/scan.err | grep '"chain": 1' | tee gpurun_out/r2x/scan_ws.jsonl
AGC_SCAN_WS=0 timeout 300 python scripts/r2_scan_bench.py 28 2>> gpurun_out/r2x/scan.err | grep '"chain": 1' | tee gpurun_out/r2x/scan_pipe.jsonl
timeout 200 python scripts/one_scan.py 28 > gpurun_out/r2x/plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_seg_scan_ws -s 2 -c 1 -o gpurun_out/r2x/prof_scan_ws python scripts/one_scan.py 28 > gpurun_out/r2x/ncu_scan.log 2>&1
tail -2 gpurun_out/r2x/ncu_scan.log
