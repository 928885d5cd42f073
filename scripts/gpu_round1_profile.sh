set -x
python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=5 --timeout 1200 -p no:cacheprovider 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench5.json 2> gpurun_out/bench5.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench5.json"))
print(d["value"], d["roofline"]["frac"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["clocks"])
for s in d["sweep"]: print(s)
PY
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16"
$B > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_G1e5.csv $B > gpurun_out/ncu_a.log 2>&1
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_cache -s 3 -c 1 -o gpurun_out/prof_cache_G1e5 $B > gpurun_out/ncu_b.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --groups 100"
$B2 > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -o gpurun_out/prof_fast_G100 $B2 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/ | tail -12
