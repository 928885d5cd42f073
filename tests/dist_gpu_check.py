"""Multi-GPU parity check, run under torchrun on >= 2 GPUs (not collected by pytest: needs several devices):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_gpu_check.py

Every rank holds a contiguous shard; aggregate_sharded must reproduce the single-GPU aggregate of the
concatenated data (bit-exact for integer / index / min / max results, 1e-12 for f64 sums)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from numpy_groupies_b200 import aggregate                     # noqa: E402
from numpy_groupies_b200.distributed import aggregate_sharded  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    rng = np.random.default_rng(7)
    n, groups = 1 << 21, 5000
    gidx = rng.integers(0, groups - 50, n)
    a64 = rng.standard_normal(n)
    a64[rng.random(n) < 0.05] = np.nan
    a32 = rng.random(n).astype(np.float32)
    ai = rng.integers(-1000, 1000, n)
    bounds = np.linspace(0, n, world + 1).astype(int)
    lo, hi = bounds[rank], bounds[rank + 1]
    checked = 0
    for func, arr, kw in (("sum", a32, {}), ("mean", a32, {}), ("max", a32, {}), ("min", a32, {}), ("len", a32, {}),
                          ("nansum", a64, {}), ("nanmean", a64, {}), ("nanvar", a64, dict(ddof=1)), ("nanstd", a64, dict(ddof=1)),
                          ("nanmax", a64, {}), ("nanargmax", a64, dict(fill_value=-1)), ("argmin", ai, dict(fill_value=-1)),
                          ("first", ai, dict(fill_value=-7)), ("last", ai, dict(fill_value=-7)), ("sum", ai, dict(fill_value=3)),
                          ("prod", 1 + a32 * 1e-3, {}), ("any", ai > 990, {}), ("all", ai > -990, dict(fill_value=1)),
                          ("nanlen", a64, {}), ("sumofsquares", a32, {})):
        full = aggregate(torch.from_numpy(gidx).cuda(), torch.from_numpy(arr).cuda(), func=func, size=groups, **kw).cpu().numpy()
        part = aggregate_sharded(torch.from_numpy(gidx[lo:hi]).cuda(), torch.from_numpy(arr[lo:hi]).cuda(), func=func,
                                 size=groups, **kw).cpu().numpy()
        assert part.dtype == full.dtype and part.shape == full.shape, (func, part.dtype, full.dtype)
        if full.dtype.kind == "f":
            scale = float(np.nanmax(np.abs(full))) if full.size else 1.0      # atol guards sums that cancel to ~0 (atomic order differs)
            np.testing.assert_allclose(part, full, rtol=1e-12 if arr.dtype == np.float64 else 1e-6,
                                       atol=(1e-13 if arr.dtype == np.float64 else 1e-6) * scale, equal_nan=True, err_msg=func)
        else:
            np.testing.assert_array_equal(part, full, err_msg=func)
        checked += 1
    # ---- deterministic fixed-order mode over shards: per-rank tables from the fixed-order kernel, folded in rank order (var / std:
    #      second pass against the global mean).  Same bits on every run, and the single-GPU result within f64 rounding ----------
    a64c = np.random.default_rng(11).standard_normal(n) + 3.0          # (a64 carries NaNs: only the nan-function sees numbers there)
    for func, arr, kw in (("sum", a64c, {}), ("mean", a64c, {}), ("var", a64c, dict(ddof=1)), ("std", a64c, {}), ("nanvar", a64, {}), ("var", a32, {})):
        full = aggregate(torch.from_numpy(gidx).cuda(), torch.from_numpy(arr).cuda(), func=func, size=groups, deterministic=True, **kw).cpu().numpy()
        parts = [aggregate_sharded(torch.from_numpy(gidx[lo:hi]).cuda(), torch.from_numpy(arr[lo:hi]).cuda(), func=func, size=groups,
                                   deterministic=True, **kw).cpu().numpy() for _ in range(2)]
        assert parts[0].tobytes() == parts[1].tobytes(), f"deterministic {func}: two runs differ"
        np.testing.assert_allclose(parts[0], full, rtol=1e-11 if arr.dtype == np.float64 else 1e-5, equal_nan=True, err_msg=f"deterministic {func}")
        checked += 1
    # tables larger than one SM's shared memory: one-word sums, split mean, hot-key cache, run kernel under ACC_ONLY
    groups2 = 120000
    for dist_name in ("uniform", "sorted", "zipf"):
        if dist_name == "uniform":
            g2 = rng.integers(0, groups2, n)
        elif dist_name == "sorted":
            g2 = np.sort(rng.integers(0, groups2, n))
        else:
            w = 1.0 / np.arange(1, groups2 + 1) ** 1.1
            g2 = rng.permutation(groups2)[np.minimum(np.searchsorted(np.cumsum(w) / w.sum(), rng.random(n)), groups2 - 1)]
        for func in ("sum", "mean", "max", "len", "var"):
            full = aggregate(torch.from_numpy(g2).cuda(), torch.from_numpy(a32).cuda(), func=func, size=groups2).cpu().numpy()
            part = aggregate_sharded(torch.from_numpy(g2[lo:hi]).cuda(), torch.from_numpy(a32[lo:hi]).cuda(), func=func, size=groups2).cpu().numpy()
            assert part.dtype == full.dtype and part.shape == full.shape, (func, part.dtype, full.dtype)
            if func in ("max", "len"):
                np.testing.assert_array_equal(part, full, err_msg=f"{func} {dist_name}")
            else:
                np.testing.assert_allclose(part, full, rtol=1e-5, atol=1e-6, err_msg=f"{func} {dist_name}")
            checked += 1
    # ---- N -> N functions: local scan + cross-GPU carry exchange; every rank returns ITS shard ---------------------
    big = rng.integers(-2 ** 40, 2 ** 40, n)
    small = rng.integers(-3, 4, n)                              # products stay interesting without dying at 0 too often
    small[small == 0] = 1
    for labels_name, gl in (("unsorted", gidx), ("sorted", np.sort(gidx))):
        for func, arr in (("cumsum", big), ("cummax", big), ("cummin", big), ("cumprod", small), ("cumsum", a32), ("nancumsum", a64),
                          ("cummax", a32), ("cumsum", ai.astype(np.int32))):
            full = aggregate(torch.from_numpy(gl).cuda(), torch.from_numpy(arr).cuda(), func=func, size=groups).cpu().numpy()
            part = aggregate_sharded(torch.from_numpy(gl[lo:hi]).cuda(), torch.from_numpy(arr[lo:hi]).cuda(), func=func, size=groups).cpu().numpy()
            assert part.dtype == full.dtype and part.shape == (hi - lo,), (func, part.dtype, part.shape)
            if full.dtype.kind == "f":
                np.testing.assert_allclose(part, full[lo:hi], rtol=1e-12 if arr.dtype == np.float64 else 1e-5,
                                           atol=1e-9 if arr.dtype == np.float64 else 1e-3, equal_nan=True, err_msg=f"{func} {labels_name}")
            else:
                np.testing.assert_array_equal(part, full[lo:hi], err_msg=f"{func} {labels_name}")          # bit-exact incl. wraparound
            checked += 1
    # ---- Form 4 (2-D group_idx) shards along N like Form 1; Form 3 on the last axis shards along the outer dimension ----
    g4 = rng.integers(0, 60, (2, n))
    for func in ("sum", "mean", "max", "argmax", "first"):
        full = aggregate(torch.from_numpy(g4).cuda(), torch.from_numpy(a64).cuda(), func="nan" + func if func in ("sum", "mean", "max", "argmax") else func,
                         size=(60, 60)).cpu().numpy()
        part = aggregate_sharded(torch.from_numpy(np.ascontiguousarray(g4[:, lo:hi])).cuda(), torch.from_numpy(a64[lo:hi]).cuda(),
                                 func="nan" + func if func in ("sum", "mean", "max", "argmax") else func, size=(60, 60)).cpu().numpy()
        assert part.shape == full.shape == (60, 60)
        if full.dtype.kind == "f":
            np.testing.assert_allclose(part, full, rtol=1e-12, atol=1e-12, equal_nan=True, err_msg=f"form4 {func}")
        else:
            np.testing.assert_array_equal(part, full, err_msg=f"form4 {func}")
        checked += 1
    rows = 64 * world
    a3 = rng.random((rows, 3000)).astype(np.float32)
    g3 = rng.integers(0, 40, 3000)
    r0, r1 = rank * 64, (rank + 1) * 64
    for func in ("sum", "max", "mean", "cumsum"):
        full = aggregate(torch.from_numpy(g3).cuda(), torch.from_numpy(a3).cuda(), func=func, size=40, axis=1).cpu().numpy()
        part = aggregate_sharded(torch.from_numpy(g3).cuda(), torch.from_numpy(a3[r0:r1]).cuda(), func=func, size=40, axis=1).cpu().numpy()
        if func == "cumsum":
            np.testing.assert_allclose(part, full[r0:r1], rtol=1e-5, err_msg="form3 cumsum")
        else:
            assert part.shape == full.shape
            np.testing.assert_allclose(part, full, rtol=1e-5, err_msg=f"form3 {func}")
        checked += 1
    # Form 3 on another axis: shards are slices along the LABELLED axis, the partial tables merge like Form 1's
    a3t = rng.random((3000, 96)).astype(np.float32)
    b3 = np.linspace(0, 3000, world + 1).astype(int)
    q0, q1 = b3[rank], b3[rank + 1]
    for func in ("sum", "mean", "max", "var"):
        full = aggregate(torch.from_numpy(g3).cuda(), torch.from_numpy(a3t).cuda(), func=func, size=40, axis=0).cpu().numpy()
        part = aggregate_sharded(torch.from_numpy(g3[q0:q1]).cuda(), torch.from_numpy(a3t[q0:q1]).cuda(), func=func, size=40, axis=0).cpu().numpy()
        assert part.shape == full.shape == (40, 96), (func, part.shape, full.shape)
        np.testing.assert_allclose(part, full, rtol=1e-5, atol=1e-6, err_msg=f"form3 axis=0 {func}")
        checked += 1
    dist.barrier()
    if rank == 0:
        print(f"dist_gpu_check ok: {checked} functions, world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
