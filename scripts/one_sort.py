"""One func='sort' call (int64 values in +-2^31, int64 labels, 10^6 groups) for a per-kernel launch list: python scripts/one_sort.py [log2n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate
n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 28)
g = torch.Generator(device="cuda").manual_seed(1)
lab = torch.randint(0, 1000000, (n,), generator=g, device="cuda")
a = torch.randint(-2 ** 31, 2 ** 31, (n,), generator=g, device="cuda")
for _ in range(2):
    out = aggregate(lab, a, func="sort", size=1000000)
torch.cuda.synchronize()
print("ok", int(out[0]))
