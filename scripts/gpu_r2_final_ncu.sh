#!/bin/bash
# --set full captures of the two headline kernels (one per call: the reports are tens of MB each)
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-r2z}
mkdir -p $O
if [ "$2" = "fast" ]; then
timeout 300 python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > $O/plain_small.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_fast -s 3 -c 1 -o $O/prof_fast python bench.py --steps 2 --warmup 1 --no-sweep --no-e2e > $O/ncu_fast.log 2>&1
tail -1 $O/ncu_fast.log
else
timeout 200 python scripts/one_scan.py 27 > $O/plain_scan.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_seg_scan_ws -s 2 -c 1 -o $O/prof_scan_ws python scripts/one_scan.py 27 > $O/ncu_scan.log 2>&1
tail -1 $O/ncu_scan.log
fi
ls -la $O/*.ncu-rep
