for u in 2 4; do echo "UPART=$u"; AGC_UPART=$u timeout 600 python scripts/tune_large.py 30 100000 2>&1 | grep '"partial"'; done
