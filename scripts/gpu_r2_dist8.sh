#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2dist8; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > $O/check_8_final.log 2>&1
tail -2 $O/check_8_final.log | cut -c1-200
for sc in weak strong; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --scaling $sc --no-sweep > $O/bench_8_${sc}_final.json 2> $O/bench_8_${sc}_final.err
  python - "$O/bench_8_${sc}_final.json" <<'PY'
import json, sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(d["scaling"], d["n_gpus"], round(d["value"],1), "Gelem/s", round(d["ms_per_step"],4), "ms/step kernel", round(d["roofline"]["kernel_ms"],4), d["verified"].get("ok"), "e2e", d.get("e2e",{}).get("value"), "check_false", d.get("value_check_false"))
except Exception as e: print("parse", e)
PY
done
timeout 300 python bench.py --steps 20 --warmup 3 --no-sweep --no-e2e > $O/bench_1_samebox.json 2> $O/bench_1.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-sweep --no-e2e --log2n 27 > $O/bench_1_2p27_samebox.json 2>> $O/bench_1.err
python - "$O" <<'PY'
import json, sys
for f in ("bench_1_samebox","bench_1_2p27_samebox"):
    d=json.loads([l for l in open(f"{sys.argv[1]}/{f}.json") if l.startswith("{")][-1])
    print(f, round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), d.get("value_check_false"))
PY
