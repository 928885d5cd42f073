#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-dist8c}; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > $O/check_8.log 2>&1
tail -3 $O/check_8.log | cut -c1-300
