#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2ar; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q --maxfail=20 > $O/pytest_gpu.log 2>&1
tail -4 $O/pytest_gpu.log | cut -c1-250
