#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2r
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q -x -k "sorted_label_scans or bad_labels or int32_labels or f32_reductions" > gpurun_out/r2r/pytest.log 2>&1
tail -3 gpurun_out/r2r/pytest.log | cut -c1-300
for cfg in "4 0" "4 1" "8 0" "8 1" "12 1" "16 1"; do
  set -- $cfg
  AGC_STAGE_THREADS=$1 AGC_NARROW_LABELS=$2 timeout 300 python scripts/r2_e2e.py 28 2>> gpurun_out/r2r/e2e.err | tee -a gpurun_out/r2r/e2e.jsonl
done
timeout 300 python scripts/r2_scan_bench.py 28 2> gpurun_out/r2r/scan.err | grep '"chain": 1' | tee gpurun_out/r2r/scan_pipe.jsonl
