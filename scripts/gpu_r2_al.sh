#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2al; mkdir -p $O
timeout 300 python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e --dist zipf > $O/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_cache -s 4 -c 1 -o $O/prof_cache python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e --dist zipf > $O/ncu.log 2>&1
tail -1 $O/ncu.log; ls -la $O
