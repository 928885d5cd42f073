"""Reductions on SORTED labels (device resident, N = 2^30 f32): python scripts/sorted_bench.py"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
n = 1 << 30
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(100)
a = torch.rand(n, generator=g, device=dev)
def timeit(fn, reps=3):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for G in (100, 100000, 10000000):
    lab = torch.sort(torch.randint(0, G, (n,), generator=g, device=dev)).values
    for func in ("sum", "mean", "max"):
        ms = timeit(lambda: aggregate(lab, a, func=func, size=G, check=False))
        aggregate(lab, a, func=func, size=G)
        print(json.dumps({"G": G, "func": func, "dist": "sorted", "ms": round(ms, 3), "gelem_s": round(n / ms * 1e-6, 1), "regime": ac.last_regime()}), flush=True)
    del lab
