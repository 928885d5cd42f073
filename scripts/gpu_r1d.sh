set -x
timeout 1200 python -m pytest tests/test_gpu_large.py -m gpu -q --maxfail=8 --timeout 600 -p no:cacheprovider -k "form3" 2>&1 | tail -15
timeout 900 python scripts/bench_configs.py 28 c5 2>&1 | cut -c1-250
