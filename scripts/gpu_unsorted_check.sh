#!/bin/bash
# unsorted N -> N / sort check: tests, then the C4 timings
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-unsorted}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_reference_suite_generic.py tests/test_gpu_parity.py -m gpu -q -x -k "n_to_n or unsorted or cum or sort" > $O/pytest.log 2>&1
tail -3 $O/pytest.log | cut -c1-250
timeout 600 python bench.py --config C4 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); [print('  ',p['func'],p['sorted'],p['ms']) for p in d['points']]"
