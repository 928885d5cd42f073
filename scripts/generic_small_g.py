"""How the general (global-table) kernel behaves at small group counts, N = 2^28 device resident."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
n = 1 << 28
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(100)
a32 = torch.rand(n, generator=g, device=dev)
a64 = a32.double()
ai = torch.randint(-1000, 1000, (n,), generator=g, device=dev)
def timeit(fn, reps=2):
    fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for G in (100, 10000):
    lab = torch.randint(0, G, (n,), generator=g, device=dev)
    for func, arr in (("sum", a64), ("mean", a64), ("max", a64), ("sum", ai), ("max", ai), ("argmax", a32), ("first", a32), ("any", ai), ("prod", a32), ("var", a32), ("len", a64)):
        ms = timeit(lambda: aggregate(lab, arr, func=func, size=G, check=False))
        aggregate(lab, arr, func=func, size=G)
        print(json.dumps({"G": G, "func": func, "dtype": str(arr.dtype).split(".")[-1], "ms": round(ms, 2), "gelem_s": round(n / ms * 1e-6, 1), "regime": ac.last_regime()}), flush=True)
