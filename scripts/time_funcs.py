"""Device-resident timing of a list of functions: python scripts/time_funcs.py log2n groups[,groups..] func[,func..] [uniform|zipf|sorted]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac

log2n = int(sys.argv[1]); groups = [int(x) for x in sys.argv[2].split(",")]; funcs = sys.argv[3].split(",")
dist = sys.argv[4] if len(sys.argv) > 4 else "uniform"
n = 1 << log2n
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(100)
a = torch.rand(n, generator=gen, device=dev)
for G in groups:
    if dist == "zipf":
        u = torch.rand(n, generator=gen, device=dev, dtype=torch.float64)
        lab = (torch.exp(u * torch.log(torch.tensor(float(G), dtype=torch.float64, device=dev))) - 1).long().clamp_(0, G - 1)
    else:
        lab = torch.randint(0, G, (n,), generator=gen, device=dev)
        if dist == "sorted":
            lab = torch.sort(lab).values
    for func in funcs:
        fn = lambda: aggregate(lab, a, func=func, size=G, check=False)
        fn(); fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(7):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        aggregate(lab, a, func=func, size=G)
        ms = min(ts)
        print(json.dumps({"G": G, "dist": dist, "func": func, "ms": round(ms, 3), "gelem_s": round(n / ms * 1e-6, 1), "regime": ac.last_regime()}), flush=True)
