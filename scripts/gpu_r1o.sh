set -x
timeout 1500 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider 2>&1 | tail -3
timeout 900 python scripts/tune_large.py 30 40000,100000,10000000 2>&1 | grep '"partial"'
