# Round-1 evidence run: bench line, ncu launch list of the same command, ncu --set full of the dominant kernels.
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench8.json 2> gpurun_out/bench8.err
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16"
$B > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_G1e5.csv $B > gpurun_out/ncu_a.log 2>&1
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o gpurun_out/prof_partial_G1e5 $B > gpurun_out/ncu_b.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --groups 100"
$B2 > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o gpurun_out/prof_fast_G100 $B2 > gpurun_out/ncu_c.log 2>&1
B3="python scripts/one_call.py"
cat > /tmp/f3.py <<PY
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from numpy_groupies_b200 import aggregate
gen = torch.Generator(device="cuda").manual_seed(100)
a = torch.rand((4096, 1 << 18), generator=gen, device="cuda")
lab = torch.randint(0, 1000, (1 << 18,), generator=gen, device="cuda")
for _ in range(3):
    aggregate(lab, a, func=sys.argv[1], size=1000, axis=1, check=False)
torch.cuda.synchronize()
PY
python /tmp/f3.py sum > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_form3 -s 2 -c 1 -f -o gpurun_out/prof_form3_sum python /tmp/f3.py sum > gpurun_out/ncu_d.log 2>&1
python /tmp/f3.py max > gpurun_out/plain_e.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_form3 -s 2 -c 1 -f -o gpurun_out/prof_form3_max python /tmp/f3.py max > gpurun_out/ncu_e.log 2>&1
$B3 cumsum 28 1000000 sorted > gpurun_out/plain_f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_seg_scan_direct -s 2 -c 2 -f -o gpurun_out/prof_scan_direct $B3 cumsum 28 1000000 sorted > gpurun_out/ncu_f.log 2>&1
$B3 cumsum 28 1000000 unsorted > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_radix_scatter -s 3 -c 1 -f -o gpurun_out/prof_radix_scatter $B3 cumsum 28 1000000 unsorted > gpurun_out/ncu_g.log 2>&1
for f in gpurun_out/prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page details > ${b}_details.txt 2>/dev/null
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  ncu -i $f --page source --csv > ${b}_source.csv 2>/dev/null
  rm -f $f
done
ls -la gpurun_out/ | tail -30
