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


# ------------------------------------------------------------------------------------------------
# the packed merge: ONE all-reduce per phase.  pack_plan() (host logic) decides how every partial table of a call rides
# in one homogeneous buffer; k_pack_tables is emulated in numpy here, the collective is real (gloo, world size 2), and the
# result must equal the per-table merge the library's agc_partials prescribes.
# ------------------------------------------------------------------------------------------------
def _np_pack(tables, xforms, status, st_x, carrier):
    segs = [(t, xforms[i]) for i, t in enumerate(tables) if i in xforms] + [(status[:3], st_x)]
    out = []
    for t, xf in segs:
        v = t.astype(np.float64 if carrier == "float64" else np.int64)
        if xf in (_abi.PK_NEG, _abi.PK_NEGFLAG):
            v = -v if carrier == "float64" else ~v
        out.append(v)
    return np.concatenate(out), segs


def _np_unpack(packed, segs, carrier):
    off = 0
    for t, xf in segs:
        v = packed[off:off + t.size]
        off += t.size
        if xf in (_abi.PK_NEG, _abi.PK_NEGFLAG):
            v = -v if carrier == "float64" else ~v
        if xf in (_abi.PK_FLAG, _abi.PK_NEGFLAG):
            v = (v != 0)
        t[...] = v.astype(t.dtype)


def _fake_tables(spec, rng):
    tabs = []
    for dt, red in spec:
        if dt == "float64":
            tabs.append(rng.standard_normal(SIZE))
        elif dt == "uint8":
            tabs.append((rng.random(SIZE) < 0.5).astype(np.uint8))
        else:
            t = rng.integers(-2 ** 62, 2 ** 62, SIZE)
            t[::5] = np.iinfo(np.int64).min                    # "empty" keys must survive the bit complement
            t[1::7] = np.iinfo(np.int64).max
            if red == "sum":
                t = rng.integers(0, 1000, SIZE)
            tabs.append(t.astype(np.int64))
    return tabs


_PACK_CASES = [("sum", 0, "float64"), ("sum", _abi.F_NEED_TOUCHED, "float64"), ("sum", _abi.F_NEED_TOUCHED, "int64"), ("mean", 0, "float64"),
               ("len", 0, "float64"), ("var", _abi.F_PASS2, "float64"), ("max", 0, "float64"), ("min", 0, "int64"), ("max", 0, "int64"),
               ("minmax", _abi.F_ACC_ONLY, "int64"), ("argmin", 0, "float64"), ("first", 0, "float64"), ("last", 0, "float64"),
               ("any", 0, "float64"), ("all", 0, "float64"), ("prod", _abi.F_NEED_TOUCHED, "float64")]


def _spec_for(op, flags, a_dtype):
    lib = _abi.load()
    req = _abi.Request()
    req.op, req.flags, req.a_dtype, req.out_dtype = _abi.OP[op], flags, _abi.DT[a_dtype], _abi.DT[a_dtype]
    req.size, req.n = SIZE, 10
    dummy = C.c_int64(0)
    req.status = C.cast(C.pointer(dummy), C.c_void_p)
    req.index.labels = C.cast(C.pointer(dummy), C.c_void_p)
    req.index.dtype = _abi.DT["int64"]
    req.out = C.cast(C.pointer(dummy), C.c_void_p)
    parts = (_abi.Partial * 6)()
    k = lib.agc_partials(C.byref(req), parts, 6)
    assert k > 0, _abi.last_error()
    return [(parts[i].dtype, parts[i].reduce) for i in range(k)]


def _pack_worker(rank, world, port, q):
    from numpy_groupies_b200.distributed import pack_plan
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    try:
        bad = []
        for ci, (op, flags, a_dtype) in enumerate(_PACK_CASES):
            codes = _spec_for(op, flags, a_dtype)
            spec = [(_abi.DT_NAME[d], _abi.REDUCE_NAMES[r]) for d, r in codes]
            rng = np.random.default_rng(1000 * ci + rank)
            tabs = _fake_tables(spec, rng)
            status = np.array([rank == 1 and ci % 2 == 0, 0, 0, 55, 0, 0, 0, 0], dtype=np.int32)
            # reference: one collective per table + one for the status
            ref = [torch.from_numpy(t.copy()) for t in tabs]
            merge_tables([(t, red) for t, (_, red) in zip(ref, spec)])
            st_ref = torch.from_numpy(status.copy())
            dist.all_reduce(st_ref, op=dist.ReduceOp.MAX)
            # packed: ONE collective (products ride alone)
            carrier, red, xforms, st_x, product = pack_plan(codes)
            got = [t.copy() for t in tabs]
            st_got = status.copy()
            for i in product:
                tp = torch.from_numpy(got[i]); dist.all_reduce(tp, op=dist.ReduceOp.PRODUCT)
            cname = _abi.DT_NAME[carrier]
            packed, segs = _np_pack(got, xforms, st_got, st_x, cname)
            tpk = torch.from_numpy(packed)
            dist.all_reduce(tpk, op=ops[red])
            _np_unpack(tpk.numpy(), segs, cname)
            for i, (t, r_) in enumerate(zip(got, ref)):
                r_np = r_.numpy()
                same = np.allclose(t, r_np, rtol=1e-12, atol=0) if t.dtype == np.float64 else np.array_equal(t != 0 if spec[i][0] == "uint8" else t, r_np != 0 if spec[i][0] == "uint8" else r_np)
                if not same:
                    bad.append((op, a_dtype, i, spec[i]))
            if bool(st_got[0]) != bool(st_ref.numpy()[0]) or st_got[3] != 55:
                bad.append((op, a_dtype, "status", st_got.tolist()))
        if rank == 0:
            q.put(bad)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_packed_merge_is_one_collective_and_equals_per_table_merge():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_pack_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    bad = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert bad == []


def test_partials_drop_unread_tables():
    """The "touched" flags only travel when the finalise step reads them (round-1 merged 100 KB of them for nothing)."""
    f64, i64, u8 = _abi.DT["float64"], _abi.DT["int64"], _abi.DT["uint8"]
    assert _spec_for("sum", 0, "float64") == [(f64, 0)]
    assert _spec_for("sum", _abi.F_NEED_TOUCHED, "float64") == [(f64, 0), (u8, 1)]
    assert _spec_for("max", 0, "float64") == [(i64, 1)]
    assert _spec_for("max", 0, "int64") == [(i64, 1), (u8, 1)]
