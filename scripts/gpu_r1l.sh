timeout 900 python scripts/tune_large.py 30 100000,10000000 2>&1 | grep '"partial"'
