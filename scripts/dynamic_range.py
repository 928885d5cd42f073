"""Wide dynamic range (log-uniform over 2^-20 .. 2^20): how much of the data falls out of the exact fixed-point window and what that costs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
gen = torch.Generator(device="cuda").manual_seed(1)
def t(lab, a, func, G):
    for _ in range(2): aggregate(lab, a, func=func, size=G, check=False)
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); aggregate(lab, a, func=func, size=G, check=False); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    aggregate(lab, a, func=func, size=G)
    return [min(ts), ac.last_regime()]
u = torch.rand(n, generator=gen, device="cuda")
for span in (8, 20, 40):
    a = torch.exp2((u - 0.5) * span)
    for G in (100, 10000, 40000, 100000):
        lab = torch.randint(0, G, (n,), generator=gen, device="cuda")
        print(json.dumps({"binades": span, "G": G, "sum": t(lab, a, "sum", G), "mean": t(lab, a, "mean", G), "rand_sum": t(lab, u, "sum", G)}), flush=True)
