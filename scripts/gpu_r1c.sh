set -x
timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=8 --timeout 600 -p no:cacheprovider -k "large_table or probe or coverage" 2>&1 | tail -15
timeout 900 python scripts/tune_large.py 30 100000 2>&1 | tee gpurun_out/tune_large.log
