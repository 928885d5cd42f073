set -x
timeout 900 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=8 --timeout 600 -p no:cacheprovider -k "cluster or probe" 2>&1 | tail -25
B="python bench.py --steps 10 --warmup 3 --no-e2e --cpu-log2n 16"
timeout 600 $B > gpurun_out/bench6.json 2> gpurun_out/bench6.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench6.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["clocks"])
for s in d.get("sweep", []): print(s)
PY
tail -5 gpurun_out/bench6.err
