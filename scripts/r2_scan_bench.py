"""C4 scans on sorted labels, chained single pass against the three-kernel variant (N = 2^28 int64 values + int64 labels).
    python scripts/r2_scan_bench.py [log2n]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
n = 1 << log2n
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(100)
G = 1000000
lab = torch.sort(torch.randint(0, G, (n,), generator=g, device=dev)).values
a = torch.randint(-2 ** 31, 2 ** 31, (n,), generator=g, device=dev)
af = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
peak = 6557.1


def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


def timeit_batched(fn, k=20):
    """k calls back to back between one pair of events: the host's per-call work hides behind the GPU's"""
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(k):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


lab32 = lab.to(torch.int32)
ms32 = timeit_batched(lambda: aggregate(lab32, a, func="cumsum", size=G, check=False))
print(json.dumps({"func": "cumsum", "dtype": "torch.int64", "labels": "int32", "log2n": log2n, "ms_back_to_back": round(ms32, 3),
                  "frac_of_peak": round(n * 20 / ms32 * 1e-6 / peak, 4), "equal_to_int64_labels": bool(torch.equal(aggregate(lab32, a, func="cumsum", size=G), aggregate(lab, a, func="cumsum", size=G)))}), flush=True)
ref = {}
for chain in (0, 1):
    ac.tuning(scan_chain=chain)
    for func, vals in (("cumsum", a), ("cummax", a), ("cummin", a), ("cumprod", a), ("cumsum", af)):
        ms = timeit(lambda: aggregate(lab, vals, func=func, size=G, check=False))
        msb = timeit_batched(lambda: aggregate(lab, vals, func=func, size=G, check=False))
        out = aggregate(lab, vals, func=func, size=G)
        key = (func, str(vals.dtype))
        same = None
        if chain == 0:
            ref[key] = out
        else:
            same = bool(torch.equal(out, ref[key])) if vals.dtype != torch.float64 else bool(torch.allclose(out, ref[key], rtol=1e-12, atol=0))
        print(json.dumps({"func": func, "dtype": str(vals.dtype), "chain": chain, "log2n": log2n, "ms": round(ms, 3), "ms_back_to_back": round(msb, 3),
                          "frac_of_peak": round(n * 24 / ms * 1e-6 / peak, 4), "equal_to_three_kernel": same}), flush=True)
