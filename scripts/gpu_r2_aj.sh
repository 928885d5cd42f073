#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2aj; mkdir -p $O
for f in 0 32 64; do
  echo "== L2 fetch $f"
  AGC_L2_FETCH=$f timeout 600 python bench.py --config C4 --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); [print('  ',p['func'],p['sorted'],p['ms']) for p in d['points']]"
  AGC_L2_FETCH=$f timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  default', round(d['value'],1)); [print('  ',s['func'],s['groups'],s['dist'],s['ms']) for s in d['sweep'] if s['groups']!=100]"
done
