set -x
B="python bench.py --steps 2 --warmup 3 --no-sweep --no-e2e --cpu-log2n 16 --log2n 28 --func ${FUNC:-sum}"
timeout 300 $B > gpurun_out/plain_cl.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cluster -s 3 -c 1 -f -o gpurun_out/prof_cluster_${FUNC:-sum} $B > gpurun_out/ncu_cl.log 2>&1
tail -3 gpurun_out/ncu_cl.log
ls -la gpurun_out/*.ncu-rep
