#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2k
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2k/pytest_gpu.log 2>&1
tail -4 gpurun_out/r2k/pytest_gpu.log | cut -c1-220
timeout 300 python scripts/r2_scan_bench.py 28 2> gpurun_out/r2k/scan.err | grep '"chain": 1' > gpurun_out/r2k/scan.jsonl; cat gpurun_out/r2k/scan.jsonl
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2k/bench.json 2> gpurun_out/r2k/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2k/bench.json").read().strip().splitlines()[-1])
print(round(d["value"],1), round(d["ms_per_step"],3), round(d["roofline"]["frac"],4), d["verified"])
for s in d["sweep"]:
    if s["groups"]==10000000: print("  ", s)
PY
