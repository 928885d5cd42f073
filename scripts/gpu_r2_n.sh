#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2n
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted" > gpurun_out/r2n/pytest_scan.log 2>&1
tail -5 gpurun_out/r2n/pytest_scan.log | cut -c1-300
echo "== pipe"
timeout 300 python scripts/r2_scan_bench.py 28 2> gpurun_out/r2n/scan.err | grep '"chain": 1' | tee gpurun_out/r2n/scan_pipe.jsonl
tail -3 gpurun_out/r2n/scan.err
