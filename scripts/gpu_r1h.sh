set -x
timeout 1500 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=8 --timeout 600 -p no:cacheprovider 2>&1 | tail -4
timeout 900 python scripts/tune_large.py 30 20000,40000,100000 2>&1 | grep -v cluster | grep mean | tee gpurun_out/tune_large5.log
