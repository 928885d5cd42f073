"""Per-kernel GPU times of a few calls in live mode (CUPTI activity trace via torch.profiler, no serialisation): python scripts/kernel_times.py func groups"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from numpy_groupies_b200 import aggregate
func, G = sys.argv[1], int(sys.argv[2])
n = 1 << 30
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand(n, generator=gen, device="cuda"); lab = torch.randint(0, G, (n,), generator=gen, device="cuda")
for _ in range(3): aggregate(lab, a, func=func, size=G, check=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): aggregate(lab, a, func=func, size=G, check=False)
    torch.cuda.synchronize()
for e in prof.events():
    if "k_" in e.name or "agc" in e.name:
        print(f"{e.name[:70]:70s} {e.device_time:10.1f} us")
