#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ac; mkdir -p $O
for cfg in "8 0 0" "8 1 1" "8 1 0" "12 1 1" "6 1 1" "8 0 1"; do
  set -- $cfg
  echo "threads $1 narrow_pinned $2 side_copy $3"
  AGC_STAGE_THREADS=$1 AGC_NARROW_PINNED=$2 AGC_SIDE_COPY=$3 timeout 300 python scripts/r2_e2e.py 28 2>> $O/e2e.err | tee -a $O/e2e_$1_$2_$3.jsonl
done
