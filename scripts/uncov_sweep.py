"""Which table size a partial-table kernel should take: 195 KB (full streaming rate) or 226 KB (more groups covered)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
n = 1 << 30
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(100)
a = torch.rand(n, generator=gen, device=dev)
def timeit(fn, reps=5):
    fn(); fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)
for G in (55000, 60000, 65000, 70000, 80000, 90000, 100000, 110000, 125000, 150000, 200000):
    lab = torch.randint(0, G, (n,), generator=gen, device=dev)
    for func in ("sum", "max", "len"):
        row = {"G": G, "func": func}
        for pct in (0, 100):                     # 0: always the 226 KB table once the 195 KB one does not cover everything; 100: always 195 KB
            ac.tuning(uncov_pct=pct)
            row["small_table" if pct else "large_table"] = round(timeit(lambda: aggregate(lab, a, func=func, size=G, check=False)), 3)
        print(json.dumps(row), flush=True)
ac.tuning(uncov_pct=10)
