#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2i
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2i/pytest_gpu.log 2>&1
tail -4 gpurun_out/r2i/pytest_gpu.log | cut -c1-220
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2i/bench.json 2> gpurun_out/r2i/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2i/bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["verified"], d["e2e"])
for s in d["sweep"]: print("  ", s)
PY
tail -2 gpurun_out/r2i/bench.err
