set -x
timeout 900 python -m pytest tests -m gpu -q --maxfail=5 --timeout 300 -p no:cacheprovider -k "n_to_n or cum or golden or parity or multi_million" 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py 28 c4 2>&1 | cut -c1-200 | head -4
AGC_SCAN_TWO_PASS=1 timeout 600 python scripts/bench_configs.py 28 c4 2>&1 | cut -c1-200 | head -4
