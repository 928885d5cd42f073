"""BASELINE config 1 (N = 1e6 float64, 1000 groups, numpy in / numpy out): latency of one call, against the reference's two CPU backends.
    python scripts/c1_latency.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from numpy_groupies_b200 import aggregate
rng = np.random.default_rng(0)
n, G = 1_000_000, 1000
g = rng.integers(0, G, n); a = rng.random(n)


def best(fn, reps=20):
    fn(); fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3


out = {"cuda_numpy_in_out_ms": round(best(lambda: aggregate(g, a, func="sum", size=G)), 3)}
tg, ta = torch.from_numpy(g).cuda(), torch.from_numpy(a).cuda()
def dev():
    aggregate(tg, ta, func="sum", size=G); torch.cuda.synchronize()
out["cuda_device_resident_ms"] = round(best(dev), 3)
ref = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
if os.path.isdir(ref):
    sys.path.insert(0, ref)
    from numpy_groupies import aggregate_numpy
    out["aggregate_numpy_ms"] = round(best(lambda: aggregate_numpy.aggregate(g, a, func="sum", size=G)), 3)
    try:
        from numpy_groupies import aggregate_numba
        out["aggregate_numba_ms"] = round(best(lambda: aggregate_numba.aggregate(g, a, func="sum", size=G)), 3)
    except Exception as e:
        out["aggregate_numba_error"] = str(e)[:80]
print(json.dumps(out))
