set -x
timeout 1500 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider 2>&1 | tail -6
timeout 900 python scripts/sorted_bench.py 2>&1 | tail -10
