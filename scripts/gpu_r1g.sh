set -x
for cfg in "cumsum 28 1000000 unsorted"; do
  tag=$(echo $cfg | tr ' ' '_')
  timeout 300 python scripts/one_call.py $cfg > gpurun_out/plain_$tag.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$tag.csv python scripts/one_call.py $cfg > gpurun_out/ncu_$tag.log 2>&1
done
