"""Throughput against the shared-memory carve-out: the same kernels with table caps below / above 196 KB per SM."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac

n = 1 << 30
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(100)
a = torch.rand(n, generator=gen, device=dev)


def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


cases = [(100000, "uniform"), (100000, "zipf"), (25000, "uniform"), (52000, "uniform"), (65000, "uniform"), (120000, "uniform"), (150000, "uniform"), (1000000, "uniform"), (10000000, "zipf")]
for G, dist in cases:
    if dist == "zipf":
        u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
        lab = (torch.exp(u * torch.log(torch.tensor(float(G), dtype=torch.float64, device=dev))) - 1).long().clamp_(0, G - 1)
        del u
    else:
        lab = torch.randint(0, G, (n,), generator=gen, device=dev)
    for func in ("sum", "mean", "max", "len"):
        row = {"G": G, "dist": dist, "func": func}
        for name, s1, s2 in (("cap_226K_110K", 227 * 1024 - 1024, 110 * 1024), ("cap_195K_97K", 199552, 99200)):
            ac.tuning(smem1=s1, smem2=s2)
            ms = timeit(lambda: aggregate(lab, a, func=func, size=G, check=False))
            aggregate(lab, a, func=func, size=G)
            row[name] = [round(ms, 3), round(n / ms * 1e-6, 1), ac.last_regime()]
        print(json.dumps(row), flush=True)
    del lab
ac.tuning(smem1=199552, smem2=99200)
