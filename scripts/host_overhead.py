"""Where the host time of a small device-resident aggregate() call goes (cProfile, 2000 calls)."""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
n, G = 1 << 17, 1000
lab = torch.randint(0, G, (n,), generator=g, device=dev)
a = torch.rand(n, generator=g, device=dev)
for _ in range(50): aggregate(lab, a, func="sum", size=G, check=False)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(2000): aggregate(lab, a, func="sum", size=G, check=False)
torch.cuda.synchronize()
print("us per call (check=False):", (time.perf_counter() - t0) / 2000 * 1e6)
t0 = time.perf_counter()
for _ in range(2000): aggregate(lab, a, func="sum", size=G)
torch.cuda.synchronize()
print("us per call (check=True):", (time.perf_counter() - t0) / 2000 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): aggregate(lab, a, func="sum", size=G, check=False)
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
