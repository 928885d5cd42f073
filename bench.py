#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: Form 1 ``aggregate(group_idx, a, 'sum')`` at
N = 2^30 float32 values, int64 labels (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--groups G] [--dist uniform|zipf]
                    [--func sum] [--log2n 30] [--no-sweep] [--no-e2e]

One "step" = one public-API call ``aggregate(labels, a, func, size=G)`` on device-resident inputs
(12 B/element, 12.9 GB per GPU: larger than the 126 MB L2, so no flush is needed between steps).
Rank 0 prints ONE JSON line (contract in the task statement): value = whole-job Gelem/s, `roofline` for the
dominant kernel (CUDA events recorded inside libaggcuda around that kernel, on the launching stream),
`cpu_baseline` = the oracle (numpy restatement of aggregate_numpy, 1 core) on a bounded sample, `e2e` = the
same call with HOST inputs (pinned) so H2D + D2H are inside the timed region, `sweep` = the other
func / group-count / distribution points of configs[1] measured in the same run (3 steps each).

N > 1: launched by torchrun, one rank per GPU; every rank holds its own N-element shard (weak scaling),
partial tables are merged with one NCCL all-reduce (numpy_groupies_b200/distributed.py).
`--impl reference` times the CPU oracle on the host cores instead (rank 0 only).
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
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-log2n", type=int, default=26)
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
# CPU legs (the oracle is the checker / baseline, never the product)
# ------------------------------------------------------------------------------------------------
def cpu_time_oracle(func, groups, dist, log2n, reps=3):
    import numpy as np
    from oracle import aggregate_oracle as oracle
    n = 1 << log2n
    labels, a = make_inputs_numpy(n, groups, dist, 100)
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.aggregate(labels, a, func=func, size=groups)
        best = min(best, time.perf_counter() - t0)
    out = {"value": n / best * 1e-9, "unit": "Gelem/s", "cores": 1, "kind": "port",
           "sample": f"N=2^{log2n} of the same workload (labels {dist} over {groups} groups, f32 values), best of {reps}, "
                     f"oracle/aggregate_oracle.py (numpy restatement of aggregate_numpy; single-threaded like the reference); "
                     f"host has {os.cpu_count()} cores"}
    try:                                             # the reference's faster CPU design: one serial compiled loop
        if func in ("sum", "mean", "max"):
            oracle.serial_loop_baseline(labels[:1000], a[:1000], func, groups)      # JIT warm-up
            t0 = time.perf_counter()
            oracle.serial_loop_baseline(labels, a, func, groups)
            out["numba_style_value"] = n / (time.perf_counter() - t0) * 1e-9
    except Exception as e:                            # noqa: BLE001
        out["numba_style_error"] = str(e)[:80]
    return out


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import aggregate_oracle as oracle
    n = 1 << args.cpu_log2n
    labels, a = make_inputs_numpy(n, args.groups, args.dist, 100)
    for _ in range(max(args.warmup, 0)):
        oracle.aggregate(labels, a, func=args.func, size=args.groups)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.aggregate(labels, a, func=args.func, size=args.groups)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt * 1e-9
    cfg = workload_config(args)
    cfg["reference_sample"] = f"each timed step aggregates N=2^{args.cpu_log2n} elements drawn like the workload's (bounded CPU sample)"
    line = {"impl": "reference", "metric": metric_name(args), "value": val, "unit": "Gelem/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": val, "unit": "Gelem/s", "cores": 1, "kind": "port",
                             "sample": f"each step = N=2^{args.cpu_log2n} elements of the workload; numpy restatement of "
                                       f"aggregate_numpy (single-threaded by construction, host has {os.cpu_count()} cores)"},
            "e2e": {"value": val, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def dominant_kernel(regime):
    special = {16: "k_len16 (16-bit counters in shared memory; mean = one-word sum pass + this count pass)",
               17: "k_maxfilter (16-bit per-group filter in shared memory + exact RED.MAX for what passes)",
               18: "k_maxfilter (16-bit per-group filter in shared memory + exact RED.MIN for what passes)",
               19: "k_fast (one-word carry-forwarding sum + u32 count per group)", 60: "k_accumulate_small (per-CTA shared-memory tables, any dtype)"}
    fam = special.get(regime) or {0: "k_accumulate (global tables)", 1: "k_fast (shared-memory table; regime 15 = one-word carry-forwarding sums, partial table + global RED when size exceeds one SM)",
           2: "k_cache (per-CTA hot-key cache + global RED)", 4: "k_form3", 5: "k_runs (segmented warp reduce)"}.get(regime // 10, "?")
    return f"accumulate stage of the step: {fam}, regime {regime}"


def metric_name(args):
    return f"aggregate {args.func} throughput (Form 1, N=2^{args.log2n}/GPU f32, int64 labels, {args.groups} groups {args.dist})"


def workload_config(args, n_override=None):
    n = n_override or (1 << args.log2n)
    return {"workload": f"form1_{args.func}_f32_int64labels_N{n}_G{args.groups}_{args.dist}",
            "func": args.func, "n_per_gpu": n, "groups": args.groups, "labels": f"int64 {args.dist}", "values": "float32 rand[0,1)",
            "l2": "inputs (12 B/elem x N) exceed the 126 MB L2; no flush needed", "parallelism": f"shard-N x{args.gpus}"}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
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
    n, G = 1 << args.log2n, args.groups
    labels = make_labels_torch(n, G, args.dist, 100 + rank, device)
    a = make_values_torch(n, 100 + rank, device)

    def step():
        if world > 1:
            return aggregate_sharded(labels, a, func=args.func, size=G)
        return aggregate(labels, a, func=args.func, size=G)

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
        kernel_ms.append(lib.agc_profile_last_ms())
    ev1.record()
    barrier()
    if sampler:
        sampler.mark("t1")
    total_ms = ev0.elapsed_time(ev1)
    launches = lib.agc_launch_count() - launches0
    lib.agc_profile_enable(0)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([total_ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = n * world / (ms_per_step * 1e-3) * 1e-9

    # ---- end-to-end through the public API with HOST (pinned) inputs, on every rank -------------------------
    # each step: H2D of this rank's shard (labels + values) -> aggregate / aggregate_sharded -> D2H of the result
    e2e = None
    if not args.no_e2e:
        try:
            e2e_log2n = min(args.log2n, 28 if world == 1 else 27)
            ne = 1 << e2e_log2n
            hl = labels[:ne].cpu().pin_memory()
            ha = a[:ne].cpu().pin_memory()

            def e2e_step():
                if world > 1:                             # aggregate_sharded takes device shards: the H2D copy is the caller's
                    r = aggregate_sharded(hl.to(device, non_blocking=True), ha.to(device, non_blocking=True), func=args.func, size=G)
                else:
                    r = aggregate(hl, ha, func=args.func, size=G)
                return r.cpu()
            e2e_step(); barrier()
            reps = 3
            t0 = time.perf_counter()
            for _ in range(reps):
                e2e_step()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / reps], device=device, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dt = float(dt.item())
            e2e = {"value": ne * world / dt * 1e-9, "unit": "Gelem/s", "h2d_bytes_per_step": ne * 12 * world, "d2h_bytes_per_step": G * 4 * world,
                   "steps": reps, "n_per_gpu": ne, "note": "pinned host tensors -> aggregate() -> host result on every rank; PCIe-bound"}
            del hl, ha
        except Exception as e:                        # noqa: BLE001
            e2e = {"value": None, "error": str(e)[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    out_bytes = G * 4
    alg_bytes = n * 12 + out_bytes
    kms = statistics.mean([k for k in kernel_ms if k > 0]) if any(k > 0 for k in kernel_ms) else ms_per_step
    achieved = alg_bytes / (kms * 1e-3) * 1e-9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{args.func}_G{G}_{args.dist}_2^{args.log2n}")
        except Exception:
            traffic = None
    line = {
        "metric": metric_name(args),
        "value": value, "unit": "Gelem/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args),
        "hbm_gbs_algorithmic": value * 12.0,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel_ms": kms,
                     "algorithmic_bytes": alg_bytes, "kernel": dominant_kernel(aggregate_cuda.last_regime())},
        "gpu_launches": int(launches), "clocks": clocks,
    }

    # (the end-to-end leg ran on every rank before this point: see e2e_leg)
    if e2e is not None:
        line["e2e"] = e2e

    # ---- the other points of BASELINE config 2, same run -------------------------------------------------
    if not args.no_sweep and world == 1:
        sweep = []
        sweep_n = n
        for dname in ("uniform", "zipf"):
            for g_ in (100, 100000, 10000000):
                lab2 = labels if (dname == args.dist and g_ == G) else make_labels_torch(sweep_n, g_, dname, 100, device)
                for f_ in ("sum", "mean", "max", "min"):
                    try:
                        for _ in range(2):
                            aggregate(lab2, a, func=f_, size=g_)
                        torch.cuda.synchronize()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        for _ in range(3):
                            aggregate(lab2, a, func=f_, size=g_)
                        e1.record(); torch.cuda.synchronize()
                        ms = e0.elapsed_time(e1) / 3
                        sweep.append({"func": f_, "groups": g_, "dist": dname, "ms": round(ms, 4),
                                      "gelem_s": round(sweep_n / ms * 1e-6, 2),
                                      "frac_of_peak": round((sweep_n * 12 + g_ * 4) / ms * 1e-6 / peak, 4)})
                    except Exception as e:            # noqa: BLE001
                        sweep.append({"func": f_, "groups": g_, "dist": dname, "error": str(e)[:120]})
                del lab2
        line["sweep"] = sweep

    del labels, a
    torch.cuda.empty_cache()
    line["cpu_baseline"] = cpu_time_oracle(args.func, G, args.dist, args.cpu_log2n)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
