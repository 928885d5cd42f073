#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2y
for cfg in "0 0" "1 0" "0 200" "1 200" "1 500" "0 1000"; do
  set -- $cfg
  echo "== gpu_scope $1 sleep $2"
  AGC_DESC_GPU=$1 AGC_LB_SLEEP=$2 timeout 300 python scripts/r2_scan_bench.py 28 2>> gpurun_out/r2y/scan.err | grep '"chain": 1' | cut -c1-120 | tee -a gpurun_out/r2y/scan_ws_$1_$2.jsonl
done
