#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2d
T=ubench/team_probe
# args: C ns mode log2n groups [budget] [cap] [reps] [max_clusters] [dist]
for C in 2 4; do for M in 0 1; do for dist in 0 1 2; do timeout 40 $T $C 0 $M 22 100001 199552 0 2 0 $dist >> gpurun_out/r2d/team.jsonl 2>&1; done; done; done
timeout 40 $T 2 1 1 22 100001 199552 8 2 3 0 >> gpurun_out/r2d/team.jsonl 2>&1
for G in 60000 100000; do for C in 2 4; do for M in 0 1; do timeout 60 $T $C 0 $M 30 $G 199552 0 3 >> gpurun_out/r2d/team.jsonl 2>&1; done; done; done
timeout 60 $T 2 0 0 30 76000 199552 0 3 >> gpurun_out/r2d/team.jsonl 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/r2d/team.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print({k:d.get(k) for k in ("C","mode","log2n","groups","dist","cov","max_rel_err","count_mismatch","ms_best","frac_6557")})
PY
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > gpurun_out/r2d/pytest_gpu.log 2>&1
tail -60 gpurun_out/r2d/pytest_gpu.log | cut -c1-220
timeout 900 python bench.py --steps 5 --warmup 3 --no-sweep > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err
tail -c 3000 gpurun_out/r2d/bench.json; tail -5 gpurun_out/r2d/bench.err
