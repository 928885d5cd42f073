set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16"
$B > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_G1e5.csv $B > gpurun_out/ncu_a.log 2>&1
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o gpurun_out/prof_sum1_G1e5 $B > gpurun_out/ncu_b.log 2>&1
for f in gpurun_out/prof_*.ncu-rep; do b=${f%.ncu-rep}; ncu -i $f --page details > ${b}_details.txt 2>/dev/null; ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null; rm -f $f; done
