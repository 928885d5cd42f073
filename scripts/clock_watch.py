"""Run one function in a loop and sample SM clock / power / throttle reasons meanwhile: python scripts/clock_watch.py func groups [seconds]"""
import sys, os, subprocess, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from numpy_groupies_b200 import aggregate
func, G = sys.argv[1], int(sys.argv[2]); secs = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
n = 1 << 30
gen = torch.Generator(device="cuda").manual_seed(1)
a = torch.rand(n, generator=gen, device="cuda"); lab = torch.randint(0, G, (n,), generator=gen, device="cuda")
for _ in range(3): aggregate(lab, a, func=func, size=G, check=False)
torch.cuda.synchronize()
mon = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader", "-lms", "100"], stdout=subprocess.PIPE, text=True)
t0 = time.time(); k = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
while time.time() - t0 < secs:
    aggregate(lab, a, func=func, size=G, check=False); k += 1
    if k % 50 == 0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
mon.terminate(); out = mon.communicate()[0].strip().splitlines()
print(json.dumps({"func": func, "G": G, "calls": k, "ms_per_call": round(e0.elapsed_time(e1) / k, 3), "samples": out[2:12]}))
