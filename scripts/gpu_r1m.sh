set -x
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-log2n 20 --no-e2e > gpurun_out/bench9.json 2> gpurun_out/bench9.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench9.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["roofline"]["kernel"])
for s in d.get("sweep", []): print(s)
PY
timeout 900 python scripts/bench_configs.py 28 c3,c5 2>&1 | cut -c1-250
timeout 600 python scripts/sorted_bench.py 2>&1 | tail -9
