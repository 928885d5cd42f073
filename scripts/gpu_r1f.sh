set -x
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 --timeout 900 -p no:cacheprovider 2>&1 | tail -8
timeout 900 python scripts/bench_configs.py 28 c4 2>&1 | cut -c1-250
