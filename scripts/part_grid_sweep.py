"""CTAs of the partial-table kernel (agc_tuning key 9) against the time of sum / mean at 10^5 groups, N = 2^30.
    python scripts/part_grid_sweep.py [log2n]"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from numpy_groupies_b200 import aggregate, _abi

log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
n, G = 1 << log2n, 100_000
gen = torch.Generator(device="cuda").manual_seed(1)
g = torch.randint(0, G, (n,), device="cuda", generator=gen)
a = torch.rand(n, device="cuda", generator=gen)
lib = _abi.load()
ref = None
for grid in (0, 147, 146, 144, 142, 140, 0, 144, 0, 144):
    lib.agc_tuning(9, grid)
    row = {"part_grid": grid}
    for func in ("sum", "mean"):
        out = aggregate(g, a, func=func, size=G)
        if func == "sum":
            if ref is None:
                ref = out.clone()
            row["max_rel_dev_vs_default"] = float(((out - ref).abs() / ref.abs()).max())
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); aggregate(g, a, func=func, size=G); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        row[func + "_ms"] = round(min(ts), 4)
    print(json.dumps(row), flush=True)
lib.agc_tuning(9, 0)
