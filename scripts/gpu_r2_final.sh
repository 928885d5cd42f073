#!/bin/bash
# Round-2 evidence run: GPU test-suite, the bench lines (default, reference arm, C3-C5), scan variants, the ncu launch list of
# the bench command and --set full captures of the two headline kernels.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2z}
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log | cut -c1-220
python -c "import __graft_entry__ as e; e.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err
python - "$O" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]+"/bench.json").read().strip().splitlines()[-1])
print("value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "verified", d["verified"].get("ok"), "always", d.get("value_check_always"), "false", d.get("value_check_false"), "e2e", d.get("e2e"))
print({k: [(p.get("func"), p.get("sorted", p.get("form", "")), p.get("ms")) for p in v.get("points", [])] for k, v in d.get("other_configs", {}).items()})
for s in d.get("sweep", []): print("  ", s)
PY
tail -3 $O/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; tail -c 600 $O/bench_reference.json
for c in C3 C4 C5; do timeout 600 python bench.py --config $c --steps 5 --warmup 3 > $O/bench_$c.json 2> $O/bench_$c.err; python - "$O/bench_$c.json" <<'PY'
import json, sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["roofline"]["frac"],3)); [print("   ", p) for p in d.get("points", [])]
PY
done
timeout 300 python scripts/r2_scan_bench.py 28 2> $O/scan.err | grep '"chain": 1' > $O/scan_ws.jsonl; cat $O/scan_ws.jsonl | cut -c1-160
AGC_SCAN_WS=0 timeout 300 python scripts/r2_scan_bench.py 28 2>> $O/scan.err | grep '"chain": 1' > $O/scan_lean.jsonl
timeout 300 python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > $O/plain_small.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > $O/ncu_launches.log 2>&1
tail -1 $O/ncu_launches.log
