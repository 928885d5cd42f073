#!/bin/bash
# Round-2 closing evidence on the last code state: GPU test-suite, smoke(), the default bench line and the reference arm.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2j}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --maxfail=40 > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log | cut -c1-220
python -c "import __graft_entry__ as e; e.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python - "$O" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]+"/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "verified", d["verified"].get("ok"), "e2e", d.get("e2e"))
for s in d.get("sweep", []): print("  ", s)
PY
tail -3 $O/bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 700 $O/bench_reference.json
