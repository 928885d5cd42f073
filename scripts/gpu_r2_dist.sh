#!/bin/bash
# multi-GPU: parity of aggregate_sharded against the single-GPU call, then weak / strong scaling points
cd "${GRAFT_REPO_ROOT:-.}"
N=${NG:-2}
mkdir -p gpurun_out/r2dist
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r2dist/check_${N}.log 2>&1
tail -6 gpurun_out/r2dist/check_${N}.log | cut -c1-300
for sc in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --scaling $sc $( [ "$sc" = "strong" ] && echo --no-sweep ) > gpurun_out/r2dist/bench_${N}_${sc}.json 2> gpurun_out/r2dist/bench_${N}_${sc}.err
  tail -2 gpurun_out/r2dist/bench_${N}_${sc}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2dist/bench_${N}_${sc}.json") if l.startswith("{")][-1])
    print("$sc", d["n_gpus"], round(d["value"],1), "Gelem/s", round(d["ms_per_step"],4), "ms/step kernel", round(d["roofline"]["kernel_ms"],4), d["gpu_launches"], d["verified"], d.get("e2e",{}).get("value"))
except Exception as e: print("parse", e)
PY
done
if [ "$N" = "2" ]; then
  timeout 600 python bench.py --steps 20 --warmup 3 --no-sweep --no-e2e > gpurun_out/r2dist/bench_1.json 2> gpurun_out/r2dist/bench_1.err
  timeout 600 python bench.py --steps 20 --warmup 3 --no-sweep --no-e2e --log2n 29 > gpurun_out/r2dist/bench_1_2p29.json 2>> gpurun_out/r2dist/bench_1.err
  timeout 600 python bench.py --steps 20 --warmup 3 --no-sweep --no-e2e --log2n 27 > gpurun_out/r2dist/bench_1_2p27.json 2>> gpurun_out/r2dist/bench_1.err
  python - <<PY
import json
for f in ("bench_1","bench_1_2p29","bench_1_2p27"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r2dist/{f}.json") if l.startswith("{")][-1])
        print(f, round(d["value"],1), round(d["ms_per_step"],4), round(d["roofline"]["kernel_ms"],4), d["gpu_launches"], d["verified"])
    except Exception as e: print(f, "parse", e)
PY
fi
