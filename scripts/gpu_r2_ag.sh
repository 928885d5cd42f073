#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ag; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 3 --no-sweep --no-e2e > $O/bench.json 2> $O/bench.err
python - "$O/bench.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4), d["verified"].get("ok"), "false", d["value_check_false"], "always", d["value_check_always"], "launches", d["gpu_launches"])
PY
tail -3 $O/bench.err
timeout 600 python bench.py --config C4 --steps 5 --warmup 3 > $O/bench_C4.json 2> $O/bench_C4.err
python - "$O/bench_C4.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
[print("   ", p) for p in d.get("points", [])]
PY
AGC_SORT_RANGE=0 timeout 600 python bench.py --config C4 --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print([p for p in d['points'] if p['func']=='sort'])"
