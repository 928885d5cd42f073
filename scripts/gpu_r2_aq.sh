#!/bin/bash
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/r2aq; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_reference_suite_generic.py tests/test_gpu_parity.py tests/test_gpu_extras.py -m gpu -q -x -k "sort or narrowed or n_to_n" > $O/pytest.log 2>&1
tail -3 $O/pytest.log | cut -c1-250
timeout 600 python bench.py --config C4 --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); [print('  ',p['func'],p['sorted'],p['ms']) for p in d['points']]"
python - <<'PY'
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
from numpy_groupies_b200 import aggregate
n=1<<28; g=torch.Generator(device="cuda").manual_seed(1)
lab=torch.randint(0,1000000,(n,),generator=g,device="cuda")
for dt in (torch.float32, torch.float64, torch.int32):
    a=(torch.rand(n,generator=g,device="cuda")*1e6).to(dt)
    for _ in range(2): aggregate(lab,a,func="sort",size=1000000)
    torch.cuda.synchronize(); t=time.perf_counter(); aggregate(lab,a,func="sort",size=1000000); torch.cuda.synchronize()
    print("sort", dt, round((time.perf_counter()-t)*1e3,1), "ms")
PY
