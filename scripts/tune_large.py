"""A/B timing of the large-table strategies (device-resident inputs, CUDA events, N = 2^log2n)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
groups_list = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["100000"])]
n = 1 << log2n
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(100)
a = torch.rand(n, generator=g, device=dev)
peak = 6557.1


def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for G in groups_list:
    lab = torch.randint(0, G, (n,), generator=g, device=dev)
    for func in ("sum", "mean", "max"):
        configs = [("cache", dict(large=0)), ("partial2w", dict(large=3, sum1=0)), ("partial", dict(large=3, sum1=1))]
        for name, kw in configs:
            ac.tuning(**kw)
            try:
                ms = timeit(lambda: aggregate(lab, a, func=func, size=G, check=False))
                aggregate(lab, a, func=func, size=G)
                print(json.dumps({"G": G, "func": func, "cfg": name, "ms": round(ms, 3), "gelem_s": round(n / ms * 1e-6, 1),
                                  "frac": round(n * 12 / ms * 1e-6 / peak, 3), "regime": ac.last_regime()}), flush=True)
            except Exception as ex:
                print(json.dumps({"G": G, "func": func, "cfg": name, "error": str(ex)[:200]}), flush=True)
        ac.tuning(large=1, sum1=1)
