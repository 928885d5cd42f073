#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ap; mkdir -p $O
for cfg in "8 0" "8 1" "12 1" "6 1"; do
  set -- $cfg
  echo "threads $1 narrow_pinned $2"
  AGC_STAGE_THREADS=$1 AGC_NARROW_PINNED=$2 timeout 300 python scripts/r2_e2e.py 28 2>> $O/e2e.err | tee -a $O/e2e_nt_$1_$2.jsonl
done
