"""Multi-GPU ``aggregate``: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Every reduction of the reference is a commutative monoid over elements, so large inputs shard
contiguously along N: rank r holds elements [offset_r, offset_r + n_r) of (group_idx, a), runs the SAME
single-GPU kernels with ``AGC_F_ACC_ONLY`` -- leaving per-group partial tables in its workspace -- and the partial
tables are merged with ONE all-reduce per phase: ``agc_pack_tables`` gathers every table of the call (f64 sums, i64
counts / order-preserving keys / global element indices, u8 flags) and the status words into one homogeneous buffer
(a table whose reduction differs from the buffer's rides transformed: flags in a sum, bit-complemented values in a
min / max), the buffer is all-reduced, and unpacked in place.  ``AGC_F_FIN_ONLY`` then turns the merged tables into
the result on every rank.  var/std and arg* insert their second pass between two merges (sum,count -> global mean ->
sum of squared deviations; extreme key -> first matching index).

What shards, and how:
  * Form 1 / 2 (flat labels) and Form 4 / 5 (``group_idx`` of shape (ndim, n)): slices along N, merge as above;
  * Form 3 on the LAST axis: slices along the outer dimension -- the outputs are disjoint, the only collective is the
    all-gather that hands every rank the full result; Form 3 on any other axis: slices along the labelled axis, partial
    tables for the whole output, merge as above;
  * ``deterministic=True``: the per-rank tables come from the fixed-order kernel and are all-gathered and folded in RANK
    order instead of all-reduced (var / std: also their second pass, against the global mean);
  * N -> N functions (cumsum / cumprod / cummax / cummin and nan-forms): local scan + cross-GPU carry exchange: the
    per-group totals of every rank are all-gathered, rank r folds the totals of the ranks before it (``agc_rank_fold``)
    and adds that carry to its local scan (``agc_apply_carry``); every rank returns ITS shard of the result;
  * ``sort`` needs whole groups on one device: replicas only.
The merge helpers take the collective as a parameter, so the host logic is exercised on CPU with the gloo backend
(tests/test_dist_gloo.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from . import aggregate_cuda as ac

_TORCH_DT = {_abi.DT["float64"]: "float64", _abi.DT["int64"]: "int64", _abi.DT["uint8"]: "uint8"}
_SCAN_REDUCTION = {"cumsum": ("sum", 0), "cumprod": ("prod", 3), "cummax": ("max", 1), "cummin": ("min", 2)}


def shard_offsets(n_local: int, group=None):
    """(global offset of this rank's first element, total N) via one all_gather of the shard lengths."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
    every = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(every, mine, group=group)
    lens = [int(t.item()) for t in every]
    return sum(lens[:rank]), sum(lens)


def merge_tables(tables, group=None):
    """All-reduce each (tensor, reduce) pair in place; reduce in {'sum','max','min','prod'}.  (One collective per table:
    the reference recipe the packed merge below is checked against.)"""
    import torch.distributed as dist
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "prod": dist.ReduceOp.PRODUCT}
    for t, red in tables:
        if t.numel():
            dist.all_reduce(t, op=ops[red], group=group)


def pack_plan(parts, with_status=True):
    """How the partial tables of one call travel in ONE buffer.

    ``parts`` = [(dtype_code, reduce_code)] as ``agc_partials`` lists them.  Returns ``(carrier_dtype_code, reduce_name,
    [xform per table], status_xform, product_tables)``: the buffer is reduced with ``reduce_name``; a table with another
    reduction rides transformed (``_abi.PK_*``); tables that multiply (prod) cannot ride and are merged on their own."""
    F64, I64 = _abi.DT["float64"], _abi.DT["int64"]
    riders = [(i, dt, red) for i, (dt, red) in enumerate(parts) if red != 3]
    product = [i for i, (dt, red) in enumerate(parts) if red == 3]
    if not riders:
        return I64, "max", {}, _abi.PK_COPY if with_status else None, product
    main = riders[0][2]
    carrier = F64 if (main == 0 and any(dt == F64 for _, dt, _ in riders)) else I64
    xforms = {}
    for i, dt, red in riders:
        if red == main:
            xforms[i] = _abi.PK_COPY
        elif main == 0:                        # a 0/1 max-table in a sum: the number of ranks that set it
            if red != 1:
                raise NotImplementedError("only flag tables ride in a sum")
            xforms[i] = _abi.PK_FLAG
        else:                                  # max in min or min in max: order-reversing bit complement
            xforms[i] = _abi.PK_NEG
    if main == 0:
        st = _abi.PK_FLAG
    elif main == 1:
        st = _abi.PK_COPY
    else:
        st = _abi.PK_NEG
    return carrier, _abi.REDUCE_NAMES[main], xforms, (st if with_status else None), product


def _merge_packed(lib, req, ws, status, group, device, stream):
    """ONE all-reduce for every partial table the library left in ``ws`` (+ status[0:3])."""
    import torch
    import torch.distributed as dist
    parts = (_abi.Partial * 6)()
    k = lib.agc_partials(C.byref(req), parts, 6)
    if k < 0:
        raise RuntimeError(_abi.last_error())
    carrier, red, xforms, st_x, product = pack_plan([(parts[i].dtype, parts[i].reduce) for i in range(k)])
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "prod": dist.ReduceOp.PRODUCT}
    for i in product:                                            # products cannot share a buffer with anything else
        p = parts[i]
        dt = getattr(torch, _TORCH_DT[p.dtype])
        nbytes = p.count * torch.empty((), dtype=dt).element_size()
        dist.all_reduce(ws[p.offset_bytes:p.offset_bytes + nbytes].view(dt), op=ops["prod"], group=group)
    segs = (_abi.PackSeg * 8)()
    nseg, total = 0, 0
    for i, xf in xforms.items():
        p = parts[i]
        segs[nseg].table, segs[nseg].count, segs[nseg].dtype, segs[nseg].xform = ws.data_ptr() + p.offset_bytes, p.count, p.dtype, xf
        total += p.count
        nseg += 1
    segs[nseg].table, segs[nseg].count, segs[nseg].dtype, segs[nseg].xform = status.data_ptr(), 3, _abi.DT["int32"], st_x
    total += 3
    nseg += 1
    packed = torch.empty(total, dtype=torch.float64 if carrier == _abi.DT["float64"] else torch.int64, device=device)
    if lib.agc_pack_tables(segs, nseg, packed.data_ptr(), carrier, 0, stream):
        raise RuntimeError(_abi.last_error())
    dist.all_reduce(packed, op=ops[red], group=group)
    if lib.agc_pack_tables(segs, nseg, packed.data_ptr(), carrier, 1, stream):
        raise RuntimeError(_abi.last_error())


def _base_request(c, lab, vals, plan, compute_out, offset):
    req = _abi.Request()
    req.op = _abi.OP[c.base]
    a_dt = np.dtype(ac._COMPUTE_AS.get(c.a_dtype.name, c.a_dtype))
    req.a_dtype, req.out_dtype = _abi.dtype_code(a_dt), _abi.dtype_code(compute_out)
    if vals is None:
        req.a = None
        req.a_scalar_f64 = float(c.a)
        req.a_scalar_i64 = int(c.a) if not isinstance(c.a, (float, np.floating)) else 0
    else:
        req.a = vals.ptr
    req.index = ac._build_index(c, lab)
    req.n, req.size = plan.n, plan.flat_size
    req.fill_f64, req.fill_i64, _ = ac._fill_numbers(c.fill_value, compute_out)
    req.ddof = float(getattr(c, "ddof", 0))
    req.index_base = offset
    req.arg_stride, req.arg_len = plan.arg_stride, plan.arg_len
    return req


def aggregate_sharded(group_idx, a, func="sum", size=None, fill_value=0, order="C", dtype=None, axis=None, group=None, **kwargs):
    """``aggregate`` over inputs sharded across the ranks of ``group`` (see the module docstring for what shards how).

    ``group_idx`` / ``a`` are this rank's shard (CUDA tensors).  ``size`` is required (a global maximum would need an
    extra collective).  Reductions return the full result on every rank; N -> N functions return this rank's shard.
    """
    import torch
    import torch.distributed as dist
    if size is None:
        raise ValueError("aggregate_sharded needs an explicit size")
    kwargs = dict(kwargs)
    check = kwargs.get("check", True)
    c = ac._prepare(group_idx, a, func, size, fill_value, order, dtype, axis, kwargs, ac._resolve_size_device)
    if c.kind != "device":
        raise NotImplementedError("aggregate_sharded: 'array' and Python callables need whole groups on one device (replicas only)")
    if c.base == "sort":
        raise NotImplementedError("aggregate_sharded: sort needs whole groups on one device (replicas only)")
    plan = c.plan
    if plan.form == "axis":
        if plan.axis == len(plan.dims) - 1 and len(plan.dims) >= 2:
            return _sharded_axis(c, group_idx, a, func, size, fill_value, order, dtype, axis, group, kwargs)
        # any other axis: the shards are slices ALONG the labelled axis (group_idx is the matching slice), every rank holds
        # partial tables for the whole output and they merge like Form 1's
        if c.base in ac._N_TO_N or c.base in ("first", "last", "argmax", "argmin"):
            raise NotImplementedError("aggregate_sharded, Form 3 along the labelled axis: reductions without element positions only")
    if c.base in ac._N_TO_N:
        if plan.form != "flat":
            raise ValueError(f"cannot reshape array of size {plan.n} into shape {tuple(plan.out_shape)}")
        return _sharded_scan(c, group, check)
    lib = _abi.load()
    device = ac._device_of(c)
    lab = ac._labels_on_device(c, device)
    vals = ac._values_on_device(c, device)
    # global element indices only matter for first / last / arg*: everything else skips the all_gather
    needs_offset = c.base in ("first", "last", "argmax", "argmin")
    offset = shard_offsets(plan.n, group)[0] if needs_offset else 0

    compute_out = np.dtype(ac._COMPUTE_AS.get(c.out_dtype.name, c.out_dtype))
    req = _base_request(c, lab, vals, plan, compute_out, offset)
    owner_merge = c.base in ("first", "last")
    base_flags = c.flags      # NEED_TOUCHED is already set whenever fill_value differs from the op's identity

    with torch.cuda.device(device):
        out = ac._empty(plan.flat_size, compute_out, device)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        req.out, req.status = out.ptr, status.data_ptr()
        req.flags = base_flags
        ws_bytes = lib.agc_workspace_bytes(C.byref(req))
        ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
        req.workspace, req.workspace_bytes = ws.data_ptr(), ws_bytes
        stream = ac._stream_ptr(device)

        def run(flags):
            req.flags = flags
            rc = lib.agc_aggregate(C.byref(req), stream)
            if rc:
                raise RuntimeError(f"agc_aggregate failed ({rc}): {_abi.last_error()}")

        ordered = c.deterministic and c.base in ac._ORDERED_OPS and vals is not None and np.issubdtype(c.a_dtype, np.floating)
        if ordered:
            # fixed-order mode: every rank adds its own elements in element order, the per-rank tables are folded in RANK order
            # (all-gather + agc_rank_fold) instead of an all-reduce whose association order is the library's business
            order_t, offsets_t = ac._group_order(c)
            req.flags = base_flags | _abi.F_ACC_ONLY
            if lib.agc_ordered_accumulate(C.byref(req), order_t.data_ptr(), offsets_t.data_ptr(), stream):
                raise RuntimeError(_abi.last_error())
            _merge_in_rank_order(lib, req, ws, group, device, stream)
            if c.base in ("var", "std"):
                # second pass against the GLOBAL mean (sums and counts of all ranks are in T0 / T1 now): every rank's squared
                # deviations in element order, then the T2 tables folded in rank order as well
                req.flags = base_flags | _abi.F_ACC_ONLY | _abi.F_PASS2
                if lib.agc_ordered_accumulate(C.byref(req), order_t.data_ptr(), offsets_t.data_ptr(), stream):
                    raise RuntimeError(_abi.last_error())
                _merge_in_rank_order(lib, req, ws, group, device, stream, second_pass=True)
        else:
            run(base_flags | _abi.F_ACC_ONLY)
            req.flags = base_flags
            _merge_packed(lib, req, ws, status, group, device, stream)
        if c.base in ("var", "std", "argmax", "argmin") and not ordered:
            run(base_flags | _abi.F_ACC_ONLY | _abi.F_PASS2 | _abi.F_NO_INIT)
            req.flags = base_flags | _abi.F_PASS2
            _merge_packed(lib, req, ws, status, group, device, stream)
        if owner_merge:
            # only the rank that holds the winning element can read its value: finalise locally (groups owned elsewhere get
            # the fill value), then one more merge hands every rank the owner's bits
            run(base_flags | _abi.F_FIN_ONLY)
            _merge_owner_values(lib, req, plan.flat_size, group, device, stream)
        else:
            run(base_flags | _abi.F_FIN_ONLY)
        if ac._want_check(c):                                    # (labels that already passed are remembered: aggregate_cuda._VALID)
            st = status.cpu()                                    # the only host synchronisation of the call
            if int(st[_abi.ST_BADLABEL]):
                raise ValueError("group_idx contains negative indices or indices too large for size")
            ac._mark_validated(c)
            ac._LAST["regime"] = int(st[_abi.ST_REGIME])
    if c.out_dtype != compute_out:
        out = ac._convert_f16(out, True)
    return ac._finish_shape_torch(c, ac._as_user_tensor(out))


def _merge_in_rank_order(lib, req, ws, group, device, stream, second_pass=False):
    """Deterministic merge of the (sum | product, count, touched) tables -- or, after the second pass of var / std, of the
    squared-deviation table T2: all-gather, then fold rank 0, 1, 2, ... in that order on every rank (agc_rank_fold with
    rank = world) -- the association order of the floating additions is fixed."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    size = req.size
    s8 = (size * 8 + 255) // 256 * 256                          # table carve-up of the reduction workspace (aggcuda.cu reduce_layout)
    t0 = ws[0:size * 8].view(torch.float64)
    t1 = ws[s8:s8 + size * 8].view(torch.int64)
    b1 = ws[3 * s8 + (size + 255) // 256 * 256:3 * s8 + (size + 255) // 256 * 256 + size]
    red = 3 if req.op == _abi.OP["prod"] else 0
    if second_pass:
        t2 = ws[2 * s8:2 * s8 + size * 8].view(torch.float64)
        gathered = torch.empty((world, size), dtype=t2.dtype, device=device)
        dist.all_gather_into_tensor(gathered, t2, group=group)
        if lib.agc_rank_fold(gathered.data_ptr(), world, size, _abi.DT["float64"], 0, t2.data_ptr(), stream):
            raise RuntimeError(_abi.last_error())
        return
    for tab, code, rcode in ((t0, _abi.DT["float64"], red), (t1, _abi.DT["int64"], 0)):
        gathered = torch.empty((world, size), dtype=tab.dtype, device=device)
        dist.all_gather_into_tensor(gathered, tab, group=group)
        if lib.agc_rank_fold(gathered.data_ptr(), world, size, code, rcode, tab.data_ptr(), stream):
            raise RuntimeError(_abi.last_error())
    dist.all_reduce(b1, op=dist.ReduceOp.MAX, group=group)      # 0/1 flags: order cannot matter


def _merge_owner_values(lib, req, size, group, device, stream):
    """first / last: ``agc_owner_pack`` writes the value bits of the groups whose winning element lives on this shard (0
    elsewhere), an int64 SUM all-reduce hands them to every rank -- at most one rank contributes per group -- and the
    unpack stores them for every non-empty group."""
    import torch
    import torch.distributed as dist
    carrier = torch.empty(size, dtype=torch.int64, device=device)
    if lib.agc_owner_pack(C.byref(req), carrier.data_ptr(), 0, stream):
        raise RuntimeError(_abi.last_error())
    dist.all_reduce(carrier, op=dist.ReduceOp.SUM, group=group)
    if lib.agc_owner_pack(C.byref(req), carrier.data_ptr(), 1, stream):
        raise RuntimeError(_abi.last_error())


def _sharded_scan(c, group, check):
    """cumsum / cumprod / cummax / cummin over shards: local scan, all-gather of the per-group totals, exclusive fold over
    the ranks before this one, carry added in a second light pass (reference semantics: aggregate_numpy.py:244-266,
    aggregate_numba.py:218-225,487-500).  Integer results are bit-identical to the single-GPU call."""
    import torch
    import torch.distributed as dist
    lib = _abi.load()
    device = ac._device_of(c)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    plan = c.plan
    out = ac._run_device(c)                                      # this rank's scan of its own shard
    red_name, red_code = _SCAN_REDUCTION[c.base]
    lab = ac._labels_on_device(c, device)
    vals = ac._values_on_device(c, device)
    compute_out = np.dtype(ac._COMPUTE_AS.get(c.out_dtype.name, c.out_dtype))
    with torch.cuda.device(device):
        stream = ac._stream_ptr(device)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        req = _base_request(c, lab, vals, plan, compute_out, 0)
        req.op = _abi.OP[red_name]
        req.flags = (c.flags & _abi.F_NANSKIP) | _abi.F_ACC_ONLY
        req.status = status.data_ptr()
        req.out = 0
        ws_bytes = lib.agc_workspace_bytes(C.byref(req))
        ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
        req.workspace, req.workspace_bytes = ws.data_ptr(), ws_bytes
        rc = lib.agc_aggregate(C.byref(req), stream)
        if rc:
            raise RuntimeError(f"agc_aggregate failed ({rc}): {_abi.last_error()}")
        parts = (_abi.Partial * 6)()
        k = lib.agc_partials(C.byref(req), parts, 6)
        if k < 1:
            raise RuntimeError(_abi.last_error())
        t0 = parts[0]                                            # the totals table (T0) always comes first
        dt = getattr(torch, _TORCH_DT[t0.dtype])
        totals = ws[t0.offset_bytes:t0.offset_bytes + t0.count * 8].view(dt)
        gathered = torch.empty((world, plan.flat_size), dtype=dt, device=device)
        dist.all_gather_into_tensor(gathered, totals, group=group)
        carry = torch.empty(plan.flat_size, dtype=dt, device=device)
        if lib.agc_rank_fold(gathered.data_ptr(), rank, plan.flat_size, t0.dtype, red_code, carry.data_ptr(), stream):
            raise RuntimeError(_abi.last_error())
        a_dt = np.dtype(ac._COMPUTE_AS.get(c.a_dtype.name, c.a_dtype))
        if rank > 0 and lib.agc_apply_carry(_abi.OP[c.base], lab.ptr, _abi.dtype_code(lab.dtype), plan.n, plan.flat_size, out.ptr,
                                            _abi.dtype_code(np.dtype(ac._COMPUTE_AS.get(out.dtype.name, out.dtype))), _abi.dtype_code(a_dt),
                                            carry.data_ptr(), t0.dtype, status.data_ptr(), stream):
            raise RuntimeError(_abi.last_error())
    return ac._finish_shape_torch(c, ac._as_user_tensor(out))


def _sharded_axis(c, group_idx, a, func, size, fill_value, order, dtype, axis, group, kwargs):
    """Form 3.  Last axis: every rank holds a slice of the OUTER dimension, its rows are nobody else's -- local aggregate,
    then one all-gather so that every rank has the full result.  (Any other axis: aggregate_sharded merges partial tables, shards
    being slices along the labelled axis.)"""
    import torch
    import torch.distributed as dist
    plan = c.plan
    nd = len(plan.dims)
    if plan.axis != nd - 1 or nd < 2:
        raise NotImplementedError("aggregate_sharded shards Form 3 along the outer dimension: the labelled axis must be the last one")
    if order != "C":
        raise NotImplementedError("aggregate_sharded returns C-ordered Form 3 results")
    local = ac.aggregate(group_idx, a, func, size=size, fill_value=fill_value, order=order, dtype=dtype, axis=axis, **kwargs)
    if c.base in ac._N_TO_N:
        return local                                             # rows are independent: the local scan is the answer
    world = dist.get_world_size(group)
    rows = torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device)
    every = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(every, rows, group=group)
    counts = [int(t.item()) for t in every]
    most = max(counts)
    padded = local if local.shape[0] == most else torch.cat([local, local.new_zeros((most - local.shape[0],) + tuple(local.shape[1:]))])
    gathered = torch.empty((world * most,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(gathered, padded.contiguous(), group=group)
    if all(n == most for n in counts):
        return gathered
    return torch.cat([gathered[r * most:r * most + counts[r]] for r in range(world)])
