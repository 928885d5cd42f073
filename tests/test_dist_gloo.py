"""World-size-2 gloo test (CPU) of the multi-GPU host logic in numpy_groupies_b200/distributed.py:
shard offsets, and the per-table merge recipe that agc_partials prescribes (which table is reduced with
which op, the two-phase var/std and arg* merges, global element indices for first/last).  The device
kernels are replaced by numpy emulations of the partial tables they produce; the merged result must equal
the oracle on the unsharded data."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from numpy_groupies_b200 import _abi
from numpy_groupies_b200.distributed import merge_tables, shard_offsets
from oracle import aggregate_oracle as oracle

SIZE = 37


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _partials_spec(op, flags=0):
    lib = _abi.load()
    req = _abi.Request()
    req.op, req.flags, req.a_dtype, req.out_dtype = _abi.OP[op], flags, _abi.DT["float64"], _abi.DT["float64"]
    req.size, req.n = SIZE, 10
    dummy = C.c_int64(0)
    req.status = C.cast(C.pointer(dummy), C.c_void_p)
    req.index.labels = C.cast(C.pointer(dummy), C.c_void_p)
    req.index.dtype = _abi.DT["int64"]
    req.out = C.cast(C.pointer(dummy), C.c_void_p)
    parts = (_abi.Partial * 5)()
    k = lib.agc_partials(C.byref(req), parts, 5)
    assert k > 0, _abi.last_error()
    return [(_abi.DT_NAME[parts[i].dtype], _abi.REDUCE_NAMES[parts[i].reduce]) for i in range(k)]


def _key(v):                      # the device's order-preserving int64 key of a float64 (agc_common.cuh key_of)
    b = np.asarray(v, dtype=np.float64).view(np.int64)
    return b ^ ((b >> 63) & np.int64(0x7fffffffffffffff))


def _unkey(k):
    return (k ^ ((k >> 63) & np.int64(0x7fffffffffffffff))).view(np.float64)


def _worker(rank, world, port, gidx, a, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bounds = np.linspace(0, len(a), world + 1).astype(int)
        lo, hi = bounds[rank], bounds[rank + 1]
        g, x = gidx[lo:hi], a[lo:hi]
        off, total = shard_offsets(len(x))
        assert off == lo and total == len(a)
        out = {}

        # sum / mean: T0 = f64 sum (sum), T1 = i64 count (sum)
        spec = _partials_spec("mean")
        assert spec == [("float64", "sum"), ("int64", "sum")]
        t0 = torch.from_numpy(np.bincount(g, weights=x, minlength=SIZE))
        t1 = torch.from_numpy(np.bincount(g, minlength=SIZE).astype(np.int64))
        merge_tables([(t0, spec[0][1]), (t1, spec[1][1])])
        out["sum"] = t0.numpy().copy()
        with np.errstate(all="ignore"):
            out["mean"] = np.where(t1.numpy() > 0, t0.numpy() / t1.numpy(), 0.0)

        # var: second phase about the GLOBAL mean, T2 = f64 sum of squared deviations (sum)
        assert _partials_spec("var", _abi.F_PASS2) == [("float64", "sum")]
        with np.errstate(all="ignore"):
            mean = t0.numpy() / t1.numpy()
        t2 = torch.from_numpy(np.bincount(g, weights=(x - mean[g]) ** 2, minlength=SIZE))
        merge_tables([(t2, "sum")])
        cnt = t1.numpy().astype(float)
        d = np.where(cnt > 1, cnt - 1, 0)
        with np.errstate(all="ignore"):
            out["var"] = np.where(d == 0, 0.0, t2.numpy() / d)

        # max: T0 = i64 key (max); first: T1 = global index (min); argmax: key (max) then index (min)
        assert _partials_spec("max")[0] == ("int64", "max")
        assert _partials_spec("first") == [("int64", "min")]
        assert _partials_spec("argmax") == [("int64", "max")] and _partials_spec("argmax", _abi.F_PASS2) == [("int64", "min")]
        keys = np.full(SIZE, np.iinfo(np.int64).min, dtype=np.int64)
        np.maximum.at(keys, g, _key(x))
        tk = torch.from_numpy(keys)
        merge_tables([(tk, "max")])
        out["max"] = np.where(tk.numpy() == np.iinfo(np.int64).min, 0.0, _unkey(tk.numpy()))
        first = np.full(SIZE, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(first, g, np.arange(lo, hi))
        tf = torch.from_numpy(first)
        merge_tables([(tf, "min")])
        out["first_idx"] = tf.numpy().copy()
        arg = np.full(SIZE, np.iinfo(np.int64).max, dtype=np.int64)
        hit = _key(x) == tk.numpy()[g]
        np.minimum.at(arg, g[hit], np.arange(lo, hi)[hit])
        ta = torch.from_numpy(arg)
        merge_tables([(ta, "min")])
        out["argmax"] = np.where(ta.numpy() == np.iinfo(np.int64).max, -1, ta.numpy())
        if rank == 0:
            q.put(out)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_rank_merge_matches_unsharded_oracle():
    rng = np.random.default_rng(3)
    n = 5001
    gidx = rng.integers(0, SIZE - 3, n)           # a few groups stay empty
    a = rng.standard_normal(n)
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, gidx, a, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    np.testing.assert_allclose(out["sum"], oracle.aggregate(gidx, a, "sum", size=SIZE), rtol=1e-12)
    np.testing.assert_allclose(out["mean"], oracle.aggregate(gidx, a, "mean", size=SIZE), rtol=1e-12)
    np.testing.assert_allclose(out["var"], oracle.aggregate(gidx, a, "var", size=SIZE, ddof=1), rtol=1e-10)
    np.testing.assert_array_equal(out["max"], oracle.aggregate(gidx, a, "max", size=SIZE))
    np.testing.assert_array_equal(out["argmax"], oracle.aggregate(gidx, a, "argmax", size=SIZE, fill_value=-1))
    first_ref = oracle.aggregate(gidx, np.arange(n), "first", size=SIZE, fill_value=np.iinfo(np.int64).max)
    np.testing.assert_array_equal(out["first_idx"], first_ref)
