#!/bin/bash
# scan-path check: the N -> N tests, then the sorted-scan timings (int64 and int32 labels)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-scan}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_reference_suite_generic.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted or cum" > $O/pytest_scan.log 2>&1
tail -3 $O/pytest_scan.log | cut -c1-300
timeout 300 python scripts/r2_scan_bench.py 28 2> $O/scan.err | grep -E '"chain": 1|int32' | cut -c1-200 | tee $O/scan.jsonl
tail -2 $O/scan.err
