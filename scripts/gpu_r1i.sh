set -x
B3="python scripts/one_call.py"
$B3 cumsum 28 1000000 unsorted > gpurun_out/plain_g.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_window_scatter -s 1 -c 1 -f -o gpurun_out/prof_window_scatter $B3 cumsum 28 1000000 unsorted > gpurun_out/ncu_g.log 2>&1
for f in gpurun_out/prof_window_scatter.ncu-rep; do
  b=${f%.ncu-rep}
  ncu -i $f --page details > ${b}_details.txt 2>/dev/null
  ncu -i $f --page raw --csv > ${b}_raw.csv 2>/dev/null
  rm -f $f
done
