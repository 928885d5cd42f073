#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-dist2}; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py > $O/check_2.log 2>&1
tail -5 $O/check_2.log | cut -c1-300
