"""Fused statistics against separate calls (device resident, N = 2^28 f32, int64 labels): python scripts/r2_multi_bench.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_multi

n = 1 << 28
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(100)
a = torch.rand(n, generator=g, device=dev)


def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)


for G in (100, 10000, 1000000):
    lab = torch.randint(0, G, (n,), generator=g, device=dev)
    for funcs in (["min", "max"], ["sum", "mean", "len"], ["mean", "var", "std"], ["min", "max", "mean", "sum", "len"]):
        fused = timeit(lambda: aggregate_multi(lab, a, funcs, size=G))
        sep = timeit(lambda: [aggregate(lab, a, f, size=G) for f in funcs])
        print(json.dumps({"groups": G, "funcs": funcs, "fused_ms": round(fused, 3), "separate_ms": round(sep, 3)}), flush=True)
