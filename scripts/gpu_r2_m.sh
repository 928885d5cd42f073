#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2m
nproc > gpurun_out/r2m/host.txt; free -g >> gpurun_out/r2m/host.txt; lscpu | head -20 >> gpurun_out/r2m/host.txt
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted" > gpurun_out/r2m/pytest_scan.log 2>&1
tail -5 gpurun_out/r2m/pytest_scan.log | cut -c1-300
for occ in 4 5 6 8; do
  echo "== warp occ $occ"
  AGC_SCAN_OCC=$occ timeout 300 python scripts/r2_scan_bench.py 28 2> gpurun_out/r2m/scan_$occ.err | grep '"chain": 1' | tee -a gpurun_out/r2m/scan_warp_occ$occ.jsonl
done
echo "== lean occ 6"
AGC_SCAN_WARP=0 AGC_SCAN_OCC=6 timeout 300 python scripts/r2_scan_bench.py 28 2>> gpurun_out/r2m/scan.err | grep '"chain": 1' | tee gpurun_out/r2m/scan_lean_occ6.jsonl
