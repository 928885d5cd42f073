# Round-1 evidence run: bench line, ncu launch list of the same command, ncu --set full of the dominant kernels.
# Only text exports of the reports come back (gpurun_out/ is capped at 64 MiB).
set -x
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16"
$B > gpurun_out/plain_a.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_G1e5.csv $B > gpurun_out/ncu_a.log 2>&1
$B > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o gpurun_out/prof_sum1_G1e5 $B > gpurun_out/ncu_b.log 2>&1
B2="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --groups 100"
$B2 > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -f -o gpurun_out/prof_fast_G100 $B2 > gpurun_out/ncu_c.log 2>&1
B4="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --dist sorted"
$B4 > gpurun_out/plain_h.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_runs -s 3 -c 1 -f -o gpurun_out/prof_runs_G1e5_sorted $B4 > gpurun_out/ncu_h.log 2>&1
B3="python scripts/one_call.py"
$B3 cumsum 28 1000000 sorted > gpurun_out/plain_f.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_seg_scan_direct -s 2 -c 2 -f -o gpurun_out/prof_scan_direct $B3 cumsum 28 1000000 sorted > gpurun_out/ncu_f.log 2>&1
$B3 cumsum 28 1000000 unsorted > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_radix_scatter -s 3 -c 1 -f -o gpurun_out/prof_radix_scatter $B3 cumsum 28 1000000 unsorted > gpurun_out/ncu_g.log 2>&1
$B3 sum 28 100 unsorted float64 > gpurun_out/plain_i.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_accumulate_small -s 1 -c 1 -f -o gpurun_out/prof_small_f64_G100 $B3 sum 28 100 unsorted float64 > gpurun_out/ncu_i.log 2>&1
B5="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --func max"
$B5 > gpurun_out/plain_j.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_maxfilter -s 6 -c 1 -f -o gpurun_out/prof_maxfilter_G1e5 $B5 > gpurun_out/ncu_j.log 2>&1
B6="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --func mean"
$B6 > gpurun_out/plain_k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_len16 -s 3 -c 1 -f -o gpurun_out/prof_len16_G1e5 $B6 > gpurun_out/ncu_k.log 2>&1
for f in gpurun_out/prof_*.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page details > ${b}_details.txt 2>/dev/null
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  rm -f $f
done
python scripts/bench_configs.py 28 > gpurun_out/configs_final.jsonl 2>&1
python scripts/sorted_bench.py > gpurun_out/sorted_final.jsonl 2>&1
python scripts/generic_small_g.py > gpurun_out/small_g_final.jsonl 2>&1
python scripts/tune_large.py 30 20000,40000,100000,10000000 > gpurun_out/large_final.jsonl 2>&1
python scripts/skew_cliff.py 30 > gpurun_out/skew_cliff_final.jsonl 2>&1
python scripts/host_overhead.py 2>&1 | head -3 > gpurun_out/host_overhead_final.txt
ls -la gpurun_out/ | tail -40
