"""Robustness of the large-table kernels to ONE moderately hot group among uniform labels (device resident, 2^30)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
n, G = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 30, 100000
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
for share in (0.0, 0.0005, 0.002, 0.005, 0.015):
    lab = torch.randint(0, G, (n,), generator=g, device=dev)
    if share:
        m = torch.rand(n, generator=g, device=dev) < share
        lab[m] = G - 7                                  # an uncovered group
        del m
    for func in ("sum", "max"):
        row = {"share": share, "func": func}
        for name, kw in (("auto", dict(large=1)), ("cache", dict(large=0)), ("partial", dict(large=3))):
            ac.tuning(**kw)
            row[name] = round(timeit(lambda: aggregate(lab, a, func=func, size=G, check=False)), 3)
            if name == "auto":
                aggregate(lab, a, func=func, size=G); row["regime"] = ac.last_regime(); row["probe"] = ac._LAST["probe"]
        ac.tuning(large=1)
        print(json.dumps(row), flush=True)
    del lab
