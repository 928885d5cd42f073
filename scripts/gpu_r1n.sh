set -x
timeout 1500 python -m pytest tests -m gpu -q --maxfail=5 --timeout 600 -p no:cacheprovider 2>&1 | tail -12
timeout 1200 python scripts/generic_small_g.py 2>&1 | tail -24
