#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ak; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_large.py tests/test_gpu_parity.py -m gpu -q -x > $O/pytest.log 2>&1
tail -3 $O/pytest.log | cut -c1-250
for h in 1 0; do
  echo "== hot $h"
  AGC_CACHE_HOT=$h timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  default', round(d['value'],1), d['verified'].get('ok')); [print('  ',s['func'],s['groups'],s['dist'],s['ms'],s['frac_of_peak']) for s in d['sweep'] if s['dist']=='zipf']"
done
