set -x
for cfg in "cumsum 28 1000000 sorted"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 python scripts/one_call.py $cfg > gpurun_out/plain_$tag.log 2>&1 && timeout 600 ncu --set full --clock-control none -k regex:k_seg_scan_single -s 1 -c 1 -f -o gpurun_out/prof_scan_single python scripts/one_call.py $cfg > gpurun_out/ncu_$tag.log 2>&1
done
ncu -i gpurun_out/prof_scan_single.ncu-rep --page details > gpurun_out/prof_scan_single_details.txt; rm -f gpurun_out/prof_scan_single.ncu-rep
