#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out/r2dist
timeout 600 python -m pytest tests/test_gpu_extras.py -m gpu -q -x -k "narrowed" > gpurun_out/r2dist/pytest_narrow.log 2>&1; tail -3 gpurun_out/r2dist/pytest_narrow.log | cut -c1-250
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > gpurun_out/r2dist/check_2_final.log 2>&1
tail -4 gpurun_out/r2dist/check_2_final.log | cut -c1-300
for sc in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --scaling $sc --no-sweep > gpurun_out/r2dist/bench_2_${sc}_final.json 2> gpurun_out/r2dist/bench_2_${sc}_final.err
  tail -c 700 gpurun_out/r2dist/bench_2_${sc}_final.json; echo
done
