#!/bin/bash
# the driver's GPU checks: the -m gpu test-suite and smoke()
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/${1:-full}; mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1
tail -3 $O/pytest_gpu.log | cut -c1-250
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
