"""One device-resident aggregate() call (after a warm-up) for ncu launch lists: python scripts/one_call.py func log2n groups sorted|unsorted [dtype]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate
func, L, G, order = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
dt = sys.argv[5] if len(sys.argv) > 5 else "int64"
n = 1 << L
gen = torch.Generator(device="cuda").manual_seed(100)
lab = torch.randint(0, G, (n,), generator=gen, device="cuda")
if order == "sorted":
    lab = torch.sort(lab).values
a = torch.randint(-2**31, 2**31, (n,), generator=gen, device="cuda") if dt == "int64" else torch.rand(n, generator=gen, device="cuda", dtype=getattr(torch, dt))
torch.cuda.synchronize()
for _ in range(int(os.environ.get("N_CALLS", "2"))):
    out = aggregate(lab, a, func=func, size=G, check=False)
torch.cuda.synchronize()
print("ok", out.shape)
