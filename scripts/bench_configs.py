"""Device-resident timing of every BASELINE.json config (C1-C5) through the public aggregate() call.
CUDA events on torch's current stream (the stream aggregate() launches on), best of `reps` after 2 warm-ups.
Algorithmic bytes follow SURVEY.md section 8(d).  Usage: python scripts/bench_configs.py [log2n_big=28] [which=all]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate

PEAK = 6557.1
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
L = int(sys.argv[1]) if len(sys.argv) > 1 else 28
which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["c1", "c3", "c4", "c5"]
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(100)


def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def report(cfg, name, n, nbytes, fn, reps=5):
    try:
        ms = timeit(fn, reps)
        print(json.dumps({"config": cfg, "case": name, "n": n, "ms": round(ms, 4), "gelem_s": round(n / ms * 1e-6, 2),
                          "alg_gbs": round(nbytes / ms * 1e-6, 1), "frac_of_peak": round(nbytes / ms * 1e-6 / PEAK, 4)}), flush=True)
    except Exception as ex:
        print(json.dumps({"config": cfg, "case": name, "error": f"{type(ex).__name__}: {str(ex)[:200]}"}), flush=True)
    torch.cuda.empty_cache()


if "c1" in which:
    n, G = 1_000_000, 1000
    lab = torch.randint(0, G, (n,), generator=gen, device=dev)
    a = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
    report("C1", "sum f64 N=1e6 G=1000 (device resident, launch-bound)", n, n * 16 + G * 8, lambda: aggregate(lab, a, func="sum", size=G, check=False), 20)
    report("C1", "sum f64 N=1e6 G=1000 (checked: +status readback)", n, n * 16 + G * 8, lambda: aggregate(lab, a, func="sum", size=G), 20)
    ln, an = lab.cpu().numpy(), a.cpu().numpy()
    report("C1", "sum f64 N=1e6 G=1000 (numpy in -> numpy out)", n, n * 16 + G * 8, lambda: aggregate(ln, an, func="sum", size=G), 10)

if "c3" in which:
    n, G = 1 << L, 1_000_000
    lab = torch.randint(0, G, (n,), generator=gen, device=dev)
    a = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
    a[torch.rand(n, generator=gen, device=dev) < 0.1] = float("nan")
    for func, kw in (("var", dict(ddof=1)), ("std", dict(ddof=1)), ("nanvar", dict(ddof=1)), ("nanstd", dict(ddof=1)), ("nanmean", {}), ("mean", {}),
                     ("argmax", {}), ("argmin", {}), ("nanargmax", {}), ("sum", {}), ("nansum", {}), ("max", {}), ("nanmax", {})):
        report("C3", f"{func} f64 10%NaN G=1e6", n, n * 16 + G * 8, lambda: aggregate(lab, a, func=func, size=G, fill_value=float("nan") if "arg" not in func else -1, check=False, **kw))
    del lab, a

if "c4" in which:
    n, G = 1 << L, 1_000_000
    lab = torch.randint(0, G, (n,), generator=gen, device=dev)
    slab = torch.sort(lab).values
    a = torch.randint(-2**31, 2**31, (n,), generator=gen, device=dev)
    for func in ("cumsum", "cummax"):
        report("C4", f"{func} int64 sorted labels G=1e6", n, n * 24, lambda: aggregate(slab, a, func=func, size=G, check=False), 3)
        report("C4", f"{func} int64 unsorted labels G=1e6", n, n * 24, lambda: aggregate(lab, a, func=func, size=G, check=False), 3)
    report("C4", "sort int64 unsorted labels G=1e6", n, n * 24, lambda: aggregate(lab, a, func="sort", size=G, check=False), 3)
    for func in ("first", "last"):
        report("C4", f"{func} int64 unsorted labels G=1e6", n, n * 16 + G * 8, lambda: aggregate(lab, a, func=func, size=G, check=False), 3)
    del lab, slab, a

if "c5" in which:
    rows, cols, G = 4096, 1 << 18, 1000
    if L < 28:
        rows >>= (28 - L)
    a = torch.rand((rows, cols), generator=gen, device=dev)
    lab = torch.randint(0, G, (cols,), generator=gen, device=dev)
    for func in ("sum", "mean", "max"):
        report("C5", f"Form3 axis=1 {rows}x2^18 f32 {func} G={G}", rows * cols, rows * cols * 4 + cols * 8 + rows * G * 4,
               lambda: aggregate(lab, a, func=func, size=G, axis=1, check=False))
    del a, lab
    n = 1 << L
    lab2 = torch.randint(0, 2048, (2, n), generator=gen, device=dev)
    v = torch.rand(n, generator=gen, device=dev)
    for func in ("sum", "mean"):
        report("C5", f"Form4 2xN labels -> 2048x2048 f32 {func}", n, n * 20 + 2048 * 2048 * 4, lambda: aggregate(lab2, v, func=func, size=(2048, 2048), check=False))
