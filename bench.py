#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: Form 1 ``aggregate(group_idx, a, 'sum')`` at
N = 2^30 float32 values, int64 labels, 10^5 groups (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--groups G] [--dist uniform|zipf|sorted]
                    [--func sum] [--log2n 30] [--scaling weak|strong] [--config C2|C3|C4|C5] [--no-sweep] [--no-e2e]

One "step" = one public-API call ``aggregate(labels, a, func, size=G)`` on device-resident inputs
(12 B/element, 12.9 GB per GPU: larger than the 126 MB L2, so no flush is needed between steps).
Rank 0 prints ONE JSON line (contract in the task statement): value = whole-job Gelem/s, `roofline` for the
dominant kernel (CUDA events recorded inside libaggcuda around that kernel, on the launching stream),
`cpu_baseline` = the REAL reference (baseline/_ref: aggregate_numpy and aggregate_numba, both single-threaded by
construction) on a bounded sample, `e2e` = the same call with HOST numpy inputs in page-locked memory so H2D + D2H are
inside the timed region (+ the same with ordinary pageable numpy arrays), `sweep` = the other func / group-count /
distribution points of configs[1] measured in the same run, `verified` = what was checked on the very data that was
timed (count conservation, sum of sums, and under torchrun the sharded result against a single-GPU call).

N > 1: launched by torchrun, one rank per GPU.  ``--scaling weak`` (default): every rank holds its own 2^log2n-element
shard; ``--scaling strong``: 2^log2n elements in total.  Partial tables are merged with one NCCL all-reduce
(numpy_groupies_b200/distributed.py).  ``--config C3|C4|C5`` times the other BASELINE.json configurations (single GPU).
``--impl reference`` times the reference package on the host cores instead (rank 0 only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda")
    ap.add_argument("--groups", type=int, default=100000)
    ap.add_argument("--dist", default="uniform")
    ap.add_argument("--func", default="sum")
    ap.add_argument("--log2n", type=int, default=30)
    ap.add_argument("--scaling", default="weak", choices=("weak", "strong"))
    ap.add_argument("--config", default="C2", choices=("C2", "C3", "C4", "C5"))
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-log2n", type=int, default=0, help="reference arm / cpu_baseline sample size (0: chosen from --steps)")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# synthetic inputs: labels uniform or Zipf(1.1) through a fixed permutation, values rand [0,1)
# ------------------------------------------------------------------------------------------------
def make_labels_torch(n, groups, dist, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    if dist == "uniform":
        return torch.randint(0, groups, (n,), generator=g, device=device, dtype=torch.int64)
    if dist == "zipf":
        w = torch.arange(1, groups + 1, device=device, dtype=torch.float64).pow(-1.1)
        cdf = torch.cumsum(w, 0)
        cdf /= cdf[-1].clone()
        perm = torch.randperm(groups, generator=g, device=device)
        out = torch.empty(n, dtype=torch.int64, device=device)
        chunk = 1 << 26
        for s in range(0, n, chunk):
            u = torch.rand(min(chunk, n - s), generator=g, device=device, dtype=torch.float64)
            r = torch.searchsorted(cdf, u).clamp_(max=groups - 1)
            out[s:s + chunk] = perm[r]
        return out
    if dist == "sorted":
        out = torch.randint(0, groups, (n,), generator=g, device=device, dtype=torch.int64)
        return torch.sort(out).values
    raise ValueError(dist)


def make_values_torch(n, seed, device):
    import torch
    g = torch.Generator(device=device).manual_seed(seed + 7)
    return torch.rand(n, generator=g, device=device, dtype=torch.float32)


def make_inputs_numpy(n, groups, dist, seed):
    import numpy as np
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        labels = rng.integers(0, groups, n, dtype=np.int64)
    elif dist == "sorted":
        labels = np.sort(rng.integers(0, groups, n, dtype=np.int64))
    else:
        w = 1.0 / np.arange(1, groups + 1) ** 1.1
        ranks = np.searchsorted(np.cumsum(w) / w.sum(), rng.random(n))
        labels = rng.permutation(groups)[np.minimum(ranks, groups - 1)].astype(np.int64)
    return labels, rng.random(n, dtype=np.float32)


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def wait_first(self, timeout=5.0):
        t = time.time()
        while not self.rows and time.time() - t < timeout:
            time.sleep(0.02)

    def mark(self, which):
        setattr(self, which, time.time())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for (ts, r) in self.rows if self.t0 is not None and self.t0 <= ts <= (self.t1 or ts) + 0.05]
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU legs: the REAL reference (baseline/_ref, installed unmodified by scripts/install_reference.py), never the product
# ------------------------------------------------------------------------------------------------
def load_reference():
    """(aggregate_numpy.aggregate, aggregate_numba.aggregate | None, kind).  Falls back to the oracle port when
    baseline/_ref did not travel (kind says which)."""
    if os.path.exists(os.path.join(REF_DIR, "numpy_groupies", "aggregate_numpy.py")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from numpy_groupies import aggregate_numpy
        try:
            from numpy_groupies import aggregate_numba
            nb = aggregate_numba.aggregate
        except Exception:                                     # noqa: BLE001  (numba missing)
            nb = None
        return aggregate_numpy.aggregate, nb, "reference"
    from oracle import aggregate_oracle as oracle
    return oracle.aggregate, None, "port"


def time_reference(func, groups, dist, log2n, reps, warm=1):
    """Both CPU backends of the reference on one bounded sample: seconds per call (best of `reps`)."""
    agg_np, agg_nb, kind = load_reference()
    n = 1 << log2n
    labels, a = make_inputs_numpy(n, groups, dist, 100)
    out = {"kind": kind, "n": n}
    for name, fn in (("aggregate_numpy", agg_np), ("aggregate_numba", agg_nb)):
        if fn is None:
            continue
        try:
            fn(labels[:4096], a[:4096], func=func, size=groups)           # JIT / import warm-up, not timed
            for _ in range(warm):
                fn(labels, a, func=func, size=groups)
            best = float("inf")
            for _ in range(reps):
                t0 = time.perf_counter()
                fn(labels, a, func=func, size=groups)
                best = min(best, time.perf_counter() - t0)
            out[name] = n / best * 1e-9
        except Exception as e:                                # noqa: BLE001
            out[name + "_error"] = str(e)[:100]
    return out


def cpu_baseline_block(func, groups, dist, log2n, reps=3):
    r = time_reference(func, groups, dist, log2n, reps)
    vals = [r[k] for k in ("aggregate_numpy", "aggregate_numba") if k in r]
    block = {"value": max(vals) if vals else None, "unit": "Gelem/s", "cores": 1, "kind": r["kind"],
             "sample": f"N=2^{log2n} of the same workload (labels {dist} over {groups} groups, f32 values), best of {reps}; value = the faster "
                       f"backend (what numpy_groupies.aggregate selects when numba is installed); both backends are single-threaded by "
                       f"construction (no prange, np.bincount is serial), host has {os.cpu_count()} cores",
             "aggregate_numpy_value": r.get("aggregate_numpy"), "aggregate_numba_value": r.get("aggregate_numba")}
    for k in r:
        if k.endswith("_error"):
            block[k] = r[k]
    return block


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    agg_np, agg_nb, kind = load_reference()
    log2n = args.cpu_log2n or (28 if args.steps + args.warmup <= 30 else 26)
    n = 1 << log2n
    labels, a = make_inputs_numpy(n, args.groups, args.dist, 100)
    res = {}
    for name, fn in (("aggregate_numpy", agg_np), ("aggregate_numba", agg_nb)):
        if fn is None:
            continue
        fn(labels[:4096], a[:4096], func=args.func, size=args.groups)       # JIT warm-up outside every timed region
        for _ in range(max(args.warmup, 0)):
            fn(labels, a, func=args.func, size=args.groups)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn(labels, a, func=args.func, size=args.groups)
        res[name] = (time.perf_counter() - t0) / args.steps
    best_name = min(res, key=res.get)
    dt = res[best_name]
    val = n / dt * 1e-9
    cfg = workload_config(args)                              # the very dict the GPU arm prints (the sample size is in cpu_baseline.sample)
    line = {"impl": "reference", "metric": metric_name(args), "value": val, "unit": "Gelem/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "reference_sample": f"each timed step aggregates N=2^{log2n} elements drawn like the workload's (bounded CPU sample)",
            "backend": best_name,
            "backends_gelem_s": {k: n / v * 1e-9 for k, v in res.items()},
            "cpu_baseline": {"value": val, "unit": "Gelem/s", "cores": 1, "kind": kind,
                             "sample": f"each step = N=2^{log2n} elements of the workload through {best_name}.aggregate of the unmodified reference "
                                       f"(baseline/_ref); the value is the faster of its two CPU backends (single-threaded by construction, host has "
                                       f"{os.cpu_count()} cores)"},
            "e2e": {"value": val, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def dominant_kernel(regime):
    special = {16: "k_len16 (16-bit counters in shared memory; mean = one-word sum pass + this count pass)",
               17: "k_maxfilter (16-bit per-group filter in shared memory + exact RED.MAX for what passes)",
               18: "k_maxfilter (16-bit per-group filter in shared memory + exact RED.MIN for what passes)",
               19: "k_fast (one-word carry-forwarding sum + u32 count per group)", 60: "k_accumulate_small (per-CTA shared-memory tables, any dtype)",
               70: "k_team (cluster of SMs owns the table: DSMEM st.async record exchange, one-word sums)",
               71: "k_team (cluster of SMs owns the table: DSMEM st.async record exchange, one-word sums + 16-bit counts)"}
    fam = special.get(regime) or {0: "k_accumulate (global tables)", 1: "k_fast (shared-memory table; regime 15 = one-word carry-forwarding sums, partial table + global RED when size exceeds one SM)",
           2: "k_cache (per-CTA hot-key cache + global RED)", 4: "k_form3", 5: "k_runs (segmented warp reduce)"}.get(regime // 10, "?")
    return f"accumulate stage of the step: {fam}, regime {regime}"


def metric_name(args):
    per = "/GPU" if args.scaling == "weak" else " total"
    return f"aggregate {args.func} throughput (Form 1, N=2^{args.log2n}{per} f32, int64 labels, {args.groups} groups {args.dist})"


def workload_config(args, n_override=None):
    n = n_override or (1 << args.log2n)
    return {"workload": f"form1_{args.func}_f32_int64labels_N{n}_G{args.groups}_{args.dist}",
            "func": args.func, "n_per_gpu" if args.scaling == "weak" else "n_total": n, "groups": args.groups, "labels": f"int64 {args.dist}",
            "values": "float32 rand[0,1)", "l2": "inputs (12 B/elem x N) exceed the 126 MB L2; no flush needed",
            "parallelism": f"shard-N x{args.gpus}", "scaling": args.scaling}


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (single GPU): --config C3 | C4 | C5
# ------------------------------------------------------------------------------------------------
def run_other_config(args):
    import torch
    from numpy_groupies_b200 import _abi, aggregate
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib = _abi.load()
    peak, peak_src = measured_peak()
    g = torch.Generator(device=dev).manual_seed(100)

    def timed(fn, steps, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    points = []
    if args.config == "C3":
        n, G = 1 << 28, 1000000
        lab = torch.randint(0, G, (n,), generator=g, device=dev)
        a = torch.rand(n, generator=g, device=dev, dtype=torch.float64)
        a[torch.rand(n, generator=g, device=dev) < 0.1] = float("nan")
        alg = n * 16 + G * 8
        desc = "C3: N=2^28 float64 with 10% NaN, 10^6 groups"
        calls = [("nanvar", dict(ddof=1)), ("nanstd", dict(ddof=1)), ("var", dict(ddof=1)), ("std", dict(ddof=1)), ("argmax", {}), ("argmin", {}),
                 ("nanargmax", {}), ("nanmean", {}), ("mean", {}), ("sum", {}), ("nansum", {}), ("max", {}), ("nanmax", {})]
        head = "nanvar"
        run = lambda f, kw: aggregate(lab, a, func=f, size=G, **kw)
        unit_n = n
    elif args.config == "C4":
        n, G = 1 << 28, 1000000
        lab_u = torch.randint(0, G, (n,), generator=g, device=dev)
        lab_s = torch.sort(lab_u).values
        a = torch.randint(-2 ** 31, 2 ** 31, (n,), generator=g, device=dev)
        alg = n * 24
        desc = "C4: N=2^28 int64 values + int64 labels, 10^6 groups"
        calls = [("cumsum", dict(sorted=True)), ("cummax", dict(sorted=True)), ("cumsum", dict(sorted=False)), ("cummax", dict(sorted=False)),
                 ("sort", dict(sorted=False)), ("first", dict(sorted=False)), ("last", dict(sorted=False))]
        head = "cumsum"

        def run(f, kw):
            return aggregate(lab_s if kw.get("sorted") else lab_u, a, func=f, size=G)
        unit_n = n
    else:
        rows, cols, G = 4096, 1 << 18, 1000
        a = torch.rand((rows, cols), generator=g, device=dev, dtype=torch.float32)
        lab = torch.randint(0, G, (cols,), generator=g, device=dev)
        n4 = 1 << 28
        lab4 = torch.randint(0, 2048, (2, n4), generator=g, device=dev)
        a4 = torch.rand(n4, generator=g, device=dev, dtype=torch.float32)
        alg = rows * cols * 4 + cols * 8 + rows * G * 4
        desc = "C5: Form 3 axis=1 of a 4096 x 2^18 float32 array (1000 groups per row); Form 4 N=2^28 -> 2048 x 2048"
        calls = [("sum", dict(form=3)), ("mean", dict(form=3)), ("max", dict(form=3)), ("sum", dict(form=4)), ("mean", dict(form=4))]
        head = "sum"

        def run(f, kw):
            if kw["form"] == 3:
                return aggregate(lab, a, func=f, size=G, axis=1)
            return aggregate(lab4, a4, func=f, size=(2048, 2048))
        unit_n = rows * cols
    launches0 = lib.agc_launch_count()
    ms_head = None
    for f, kw in calls:
        try:
            is_head = ms_head is None and f == head
            ms = timed(lambda: run(f, kw), args.steps if is_head else 3, max(args.warmup, 3) if is_head else 2)
            if args.config == "C4":
                bytes_ = n * (16 if f in ("first", "last") else 24)
                nn = n
            elif args.config == "C5":
                bytes_ = alg if kw["form"] == 3 else n4 * 20 + 2048 * 2048 * 4
                nn = unit_n if kw["form"] == 3 else n4
            else:
                bytes_, nn = alg, unit_n
            pt = {"func": f, **{k: v for k, v in kw.items()}, "ms": round(ms, 4), "gelem_s": round(nn / ms * 1e-6, 2),
                  "frac_of_peak": round(bytes_ / ms * 1e-6 / peak, 4)}
            if is_head:
                ms_head, head_bytes, head_n = ms, bytes_, nn
            points.append(pt)
        except Exception as e:                                # noqa: BLE001
            points.append({"func": f, **kw, "error": str(e)[:120]})
    launches = lib.agc_launch_count() - launches0
    line = {"metric": f"aggregate {head} throughput ({desc})", "value": head_n / ms_head * 1e-6, "unit": "Gelem/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_head, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"C3": "f64", "C4": "int64", "C5": "f32"}[args.config], "data": "synthetic",
            "config": {"workload": desc, "l2": "inputs exceed the 126 MB L2; no flush needed"},
            "roofline": {"bound": "hbm", "achieved": head_bytes / ms_head * 1e-6, "peak": peak, "unit": "GB/s", "frac": head_bytes / ms_head * 1e-6 / peak,
                         "traffic": None, "peak_source": peak_src, "kernel_ms": ms_head, "algorithmic_bytes": head_bytes,
                         "kernel": "whole call (several kernels), CUDA events around the public API call"},
            "gpu_launches": int(launches), "points": points}
    del a
    torch.cuda.empty_cache()
    if getattr(args, "_embed", False):
        return line
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config != "C2":
        if int(os.environ.get("RANK", "0")) == 0:
            run_other_config(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from numpy_groupies_b200 import _abi, aggregate, aggregate_cuda
    from numpy_groupies_b200.distributed import aggregate_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    lib = _abi.load()
    G = args.groups
    n = (1 << args.log2n) if args.scaling == "weak" else (1 << args.log2n) // world
    labels = make_labels_torch(n, G, args.dist, 100 + rank, device)
    a = make_values_torch(n, 100 + rank, device)

    def step(lab=None, val=None, func=None, groups=None):
        lab = labels if lab is None else lab
        val = a if val is None else val
        if world > 1:
            return aggregate_sharded(lab, val, func=func or args.func, size=groups or G)
        return aggregate(lab, val, func=func or args.func, size=groups or G)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    lib.agc_profile_enable(1)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_first()
    launches0 = lib.agc_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    if sampler:
        sampler.mark("t0")
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    if sampler:
        sampler.mark("t1")
    total_ms = ev0.elapsed_time(ev1)
    # the accumulate stage of every timed step (event pairs recorded inside agc_aggregate, read back only now: the timed loop
    # itself never synchronises beyond what the public call does)
    kernel_ms = [lib.agc_profile_ms(b) for b in range(min(args.steps, 64))]
    launches = lib.agc_launch_count() - launches0
    lib.agc_profile_enable(0)
    regime = aggregate_cuda.last_regime()
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3) * 1e-9

    # ---- the same K steps with the status record read back on EVERY call (check="always": one host synchronisation per step,
    #      the host cannot prepare step k+1 while step k runs) and with no read-back at all (check=False).  `value` above is the
    #      public default: the bad-label verdict depends on the labels and `size` alone, so a CUDA label tensor that passed once
    #      (same object, same torch version counter) is not read back again -- the warm-up steps validated `labels`. -----------
    def timed_variant(check):
        try:
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                if world > 1:
                    aggregate_sharded(labels, a, func=args.func, size=G, check=check)
                else:
                    aggregate(labels, a, func=args.func, size=G, check=check)
            e1.record(); barrier()
            ta = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ta, op=dist.ReduceOp.MAX)
            return n * world / (float(ta.item()) / args.steps * 1e-3) * 1e-9
        except Exception:                                     # noqa: BLE001
            return None
    async_value = timed_variant(False)
    strict_value = timed_variant("always")

    # ---- verify what was timed, on the very data that was timed --------------------------------------------------
    verified = {}
    try:
        counts = step(func="len")
        sums = step(func="sum").double()
        n_tot = torch.tensor([float(n)], device=device, dtype=torch.float64)
        s_tot = a.double().sum().reshape(1)
        if world > 1:
            dist.all_reduce(n_tot); dist.all_reduce(s_tot)
        verified["count_conservation"] = bool(int(counts.sum().item()) == int(n_tot.item()))
        verified["sum_of_sums_rel_err"] = float(abs(sums.sum().item() - s_tot.item()) / abs(s_tot.item()))
        if world > 1:                                       # sharded == single GPU, on a 2^22-element sample of every shard
            m = min(n, 1 << 22)
            gl = [torch.empty(m, dtype=labels.dtype, device=device) for _ in range(world)]
            ga = [torch.empty(m, dtype=a.dtype, device=device) for _ in range(world)]
            dist.all_gather(gl, labels[:m].contiguous()); dist.all_gather(ga, a[:m].contiguous())
            ok = True
            for f_ in ("sum", "mean", "max", "len"):
                part = aggregate_sharded(labels[:m], a[:m], func=f_, size=G)
                full = aggregate(torch.cat(gl), torch.cat(ga), func=f_, size=G)
                if f_ in ("max", "len"):
                    ok = ok and bool(torch.equal(part, full))
                else:
                    ok = ok and bool(torch.allclose(part, full, rtol=1e-5, atol=1e-30))
            flag = torch.tensor([1 if ok else 0], device=device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            verified["sharded_equals_single_gpu"] = bool(flag.item())
        verified["ok"] = bool(verified["count_conservation"] and verified["sum_of_sums_rel_err"] < 1e-6
                              and verified.get("sharded_equals_single_gpu", True))
    except Exception as e:                                    # noqa: BLE001
        verified = {"ok": False, "error": str(e)[:200]}

    # ---- end-to-end through the public API with HOST numpy inputs, on every rank ------------------------------------
    # each step: H2D of this rank's shard (labels + values) -> aggregate / aggregate_sharded -> D2H of the result.
    # `value`: numpy arrays that live in page-locked memory (one asynchronous DMA each, the contract's "pinned host memory");
    # `pageable_numpy_value`: ordinary numpy arrays, staged through the library's pinned ring (aggregate_cuda._upload_staged)
    e2e = None
    if not args.no_e2e:
        try:
            ne = n if world == 1 else min(n, 1 << 28)        # one GPU: the full workload; under torchrun 2^28 per rank (8 ranks share one host)
            hl_t, ha_t = labels[:ne].cpu().pin_memory(), a[:ne].cpu().pin_memory()
            hl, ha = hl_t.numpy(), ha_t.numpy()               # numpy views of the pinned buffers

            def e2e_step(xl, xa):
                if world > 1:                                 # aggregate_sharded takes device shards: the H2D copy is the caller's
                    r = aggregate_sharded(torch.from_numpy(xl).to(device, non_blocking=True), torch.from_numpy(xa).to(device, non_blocking=True),
                                          func=args.func, size=G)
                    return r.cpu().numpy()
                return aggregate(xl, xa, func=args.func, size=G)          # numpy in, numpy out

            def time_e2e(xl, xa, reps):
                e2e_step(xl, xa); barrier()
                t0 = time.perf_counter()
                for _ in range(reps):
                    e2e_step(xl, xa)
                barrier()
                dt = torch.tensor([(time.perf_counter() - t0) / reps], device=device, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                return float(dt.item())
            reps = 3
            dt_pinned = time_e2e(hl, ha, reps)
            e2e = {"value": ne * world / dt_pinned * 1e-9, "unit": "Gelem/s", "h2d_bytes_per_step": ne * 12 * world, "d2h_bytes_per_step": G * 4 * world,
                   "steps": reps, "n_per_gpu": ne, "note": "numpy arrays in page-locked host memory -> aggregate() -> numpy result, every rank; PCIe-bound"}
            if world == 1:
                pl, pa = np.array(hl), np.array(ha)           # pageable copies
                dt_page = time_e2e(pl, pa, reps)
                e2e["pageable_numpy_value"] = ne / dt_page * 1e-9
                del pl, pa
            del hl, ha, hl_t, ha_t
        except Exception as e:                                # noqa: BLE001
            e2e = {"value": None, "error": str(e)[:200]}

    peak, peak_src = measured_peak()
    # ---- the other points of BASELINE config 2, same run (all ranks take part under torchrun) ------------------------
    sweep = None
    if not args.no_sweep:
        sweep = []
        for dname in ("uniform", "zipf"):
            for g_ in (100, 100000, 10000000):
                lab2 = labels if (dname == args.dist and g_ == G) else make_labels_torch(n, g_, dname, 100 + rank, device)
                for f_ in ("sum", "mean", "max", "min"):
                    try:
                        for _ in range(2):
                            step(lab2, a, f_, g_)
                        barrier()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(3):
                            step(lab2, a, f_, g_)
                        e1.record(); barrier()
                        ms_t = torch.tensor([e0.elapsed_time(e1) / 3], device=device, dtype=torch.float64)
                        if world > 1:
                            dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
                        ms = float(ms_t.item())
                        sweep.append({"func": f_, "groups": g_, "dist": dname, "ms": round(ms, 4),
                                      "gelem_s": round(n * world / ms * 1e-6, 2),
                                      "frac_of_peak": round((n * 12 + g_ * 4) / ms * 1e-6 / peak, 4)})
                    except Exception as e:            # noqa: BLE001
                        sweep.append({"func": f_, "groups": g_, "dist": dname, "error": str(e)[:120]})
                # the int32-label variant SURVEY 8(d) asks for (8 instead of 12 B/element), at the headline group count
                if g_ == 100000 and dname == "uniform" and world == 1:
                    lab32 = lab2.to(torch.int32)
                    for f_ in ("sum", "max"):
                        try:
                            for _ in range(2):
                                step(lab32, a, f_, g_)
                            barrier()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record()
                            for _ in range(3):
                                step(lab32, a, f_, g_)
                            e1.record(); barrier()
                            ms = e0.elapsed_time(e1) / 3
                            sweep.append({"func": f_, "groups": g_, "dist": dname, "labels": "int32", "ms": round(ms, 4), "gelem_s": round(n / ms * 1e-6, 2),
                                          "frac_of_peak": round((n * 8 + g_ * 4) / ms * 1e-6 / peak, 4)})
                        except Exception as e:        # noqa: BLE001
                            sweep.append({"func": f_, "groups": g_, "dist": dname, "labels": "int32", "error": str(e)[:120]})
                    del lab32
                del lab2

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    out_bytes = G * 4
    alg_bytes = n * 12 + out_bytes
    kms = statistics.mean([k for k in kernel_ms if k > 0]) if any(k > 0 for k in kernel_ms) else ms_per_step
    achieved = alg_bytes / (kms * 1e-3) * 1e-9
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{args.func}_G{G}_{args.dist}_2^{args.log2n}")
            traffic_src = "profiles/traffic.json: dram bytes of ONE launch from the committed `ncu --set full` capture, not measured in this run"
        except Exception:
            traffic = None
    line = {
        "metric": metric_name(args),
        "value": value, "unit": "Gelem/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "hbm_gbs_algorithmic": value * 12.0,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel_ms": kms,
                     "algorithmic_bytes": alg_bytes, "kernel": dominant_kernel(regime)},
        "gpu_launches": int(launches), "clocks": clocks, "verified": verified,
        "value_check_false": async_value, "value_check_always": strict_value,
        "status_check": "default check=True: the status record (bad labels) is read back until a CUDA label tensor has passed once for this size "
                        "(same tensor object and version counter), later calls stay asynchronous; value_check_always = read back on every call",
    }
    if e2e is not None:
        line["e2e"] = e2e
    if sweep is not None:
        line["sweep"] = sweep

    del labels, a
    torch.cuda.empty_cache()
    # ---- the other BASELINE.json configurations (C3 f64 / 10^6 groups, C4 N -> N and sort, C5 Forms 3 and 4), one GPU: a few
    #      steps each, so that the driver's own run of this file carries them too (`--config C3|C4|C5` prints them alone) ------
    if world == 1 and not args.no_sweep:
        others = {}
        for cfg in ("C3", "C4", "C5"):
            try:
                sub = argparse.Namespace(**vars(args))
                sub.config, sub.steps, sub.warmup, sub._embed = cfg, 3, 3, True
                o = run_other_config(sub)
                others[cfg] = {"workload": o["config"]["workload"], "points": o["points"]}
            except Exception as e:                            # noqa: BLE001
                others[cfg] = {"error": str(e)[:200]}
        line["other_configs"] = others
    try:
        line["cpu_baseline"] = cpu_baseline_block(args.func, G, args.dist, args.cpu_log2n or 26)
    except Exception as e:                                    # noqa: BLE001
        line["cpu_baseline"] = {"value": None, "error": str(e)[:200]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
