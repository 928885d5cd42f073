#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
for rep in 1 2 3; do for t in 8 12 14; do AGC_STAGE_THREADS=$t python scripts/r2_e2e.py 28 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['threads'], d['pinned_ms'], d['pageable_ms'])"; done; done
