"""End-to-end timing of aggregate() with HOST numpy inputs (page-locked and pageable), Form 1 sum, 10^5 groups.
    AGC_STAGE_THREADS=8 AGC_NARROW_LABELS=1 python scripts/r2_e2e.py [log2n] [func]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from numpy_groupies_b200 import aggregate

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
func = sys.argv[2] if len(sys.argv) > 2 else "sum"
n, G = 1 << log2n, 100000
g = torch.Generator().manual_seed(5)
hl_t = torch.randint(0, G, (n,), generator=g).pin_memory()
ha_t = torch.rand(n, generator=g).pin_memory()
hl, ha = hl_t.numpy(), ha_t.numpy()


def timeit(xl, xa, reps=3):
    aggregate(xl, xa, func=func, size=G); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = aggregate(xl, xa, func=func, size=G)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


dt_pin, r1 = timeit(hl, ha)
pl, pa = np.array(hl), np.array(ha)
dt_page, r2 = timeit(pl, pa)
ref = np.bincount(pl, weights=pa.astype(np.float64), minlength=G) if func == "sum" else None
ok = bool(np.allclose(r1, ref, rtol=1e-5)) and bool(np.array_equal(r1, r2)) if ref is not None else None
print(json.dumps({"log2n": log2n, "func": func, "threads": os.environ.get("AGC_STAGE_THREADS", "default"), "narrow": os.environ.get("AGC_NARROW_LABELS", "1"),
                  "pinned_ms": round(dt_pin * 1e3, 2), "pinned_gelem_s": round(n / dt_pin * 1e-9, 3),
                  "pageable_ms": round(dt_page * 1e3, 2), "pageable_gelem_s": round(n / dt_page * 1e-9, 3), "ok": ok, "cpus": os.cpu_count()}))
