"""Does the relative placement of the label and value arrays change the kernel time?  sum at 10^5 groups, values taken at different element offsets of one big buffer."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
n = 1 << 30
G = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
func = sys.argv[2] if len(sys.argv) > 2 else "sum"
gen = torch.Generator(device="cuda").manual_seed(1)
lab = torch.randint(0, G, (n,), generator=gen, device="cuda")
big = torch.rand(n + (1 << 24), generator=gen, device="cuda")
def t(a):
    for _ in range(2): aggregate(lab, a, func=func, size=G, check=False)
    ts = []
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); aggregate(lab, a, func=func, size=G, check=False); e1.record(); torch.cuda.synchronize()
        ts.append(round(e0.elapsed_time(e1), 3))
    return min(ts)
print("labels at", hex(lab.data_ptr()), "values base", hex(big.data_ptr()))
for off in (0, 4, 64, 1024, 4096, 1 << 14, 1 << 16, 1 << 18, 1 << 19, 1 << 20, 3 << 19, 1 << 21, 1 << 22, 1 << 23):
    a = big[off:off + n]
    print(json.dumps({"offset_bytes": off * 4, "ms": t(a)}), flush=True)
