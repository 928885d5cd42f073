set -x
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider 2>&1 | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --cpu-log2n 20 > gpurun_out/bench7.json 2> gpurun_out/bench7.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench7.json").read().strip().splitlines()[-1])
print(d["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["clocks"], d.get("e2e"))
for s in d.get("sweep", []): print(s)
PY
tail -5 gpurun_out/bench7.err
timeout 900 python scripts/bench_configs.py 28 2>&1 | cut -c1-250 | tee gpurun_out/configs_r1b.log
