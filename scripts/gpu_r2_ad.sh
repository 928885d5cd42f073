#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ad; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_reference_suite_generic.py tests/test_gpu_parity.py -m gpu -q -x -k "sorted_label_scans or n_to_n or unsorted or cum or sort" > $O/pytest_scan.log 2>&1
tail -3 $O/pytest_scan.log | cut -c1-300
timeout 600 python bench.py --config C4 --steps 5 --warmup 3 > $O/bench_C4.json 2> $O/bench_C4.err
python - "$O/bench_C4.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
[print("   ", p) for p in d.get("points", [])]
PY
AGC_UNSORT_WINDOW=0 timeout 600 python bench.py --config C4 --steps 5 --warmup 3 > $O/bench_C4_nowindow.json 2>> $O/bench_C4.err
python - "$O/bench_C4_nowindow.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
[print("  old", p) for p in d.get("points", []) if not p.get("sorted", True)]
PY
