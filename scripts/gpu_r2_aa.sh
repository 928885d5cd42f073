#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ab; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_large.py tests/test_reference_suite_generic.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted or cum" > $O/pytest_scan.log 2>&1
tail -3 $O/pytest_scan.log | cut -c1-300
for i in 1 2; do timeout 300 python scripts/r2_scan_bench.py 28 2>> $O/scan.err | grep '"chain": 1' | cut -c1-150 | tee -a $O/scan_ws_coop.jsonl; done
