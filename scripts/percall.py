"""Per-call device times of one function (CUDA events around every call): python scripts/percall.py func groups calls"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate
func, G, K = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
n = 1 << 30
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand(n, generator=gen, device="cuda"); lab = torch.randint(0, G, (n,), generator=gen, device="cuda")
ts = []
for i in range(K):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(); aggregate(lab, a, func=func, size=G, check=False); e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1), 3))
print(json.dumps({"func": func, "G": G, "blocking": os.environ.get("CUDA_LAUNCH_BLOCKING", "0"), "ms": ts}))
