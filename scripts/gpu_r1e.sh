set -x
timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=8 --timeout 600 -p no:cacheprovider 2>&1 | tail -4
timeout 900 python scripts/tune_large.py 30 100,20000,100000 2>&1 | grep -v cluster | tee gpurun_out/tune_large3.log
