"""Multi-GPU ``aggregate``: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch).

Every reduction of the reference is a commutative monoid over elements, so large inputs
shard contiguously along N: rank r holds elements [offset_r, offset_r + n_r) of (group_idx, a), runs
the SAME single-GPU kernels with ``AGC_F_ACC_ONLY`` -- leaving per-group partial tables in its
workspace -- the partial tables are combined with one all-reduce each (sum / max / min as
``agc_partials`` prescribes: counts and f64 sums add, order-preserving integer keys take max/min,
first/last/arg indices take min/max of GLOBAL element indices), and ``AGC_F_FIN_ONLY`` turns the
merged tables into the result on every rank.  var/std and arg* insert their second pass between two
merges (sum,count -> global mean -> sum of squared deviations; extreme key -> first matching index).

Only Form 1 (flat labels) is sharded; N->N functions and ``sort`` need groups co-located and are
"replicas only" (DESIGN.md).  The merge helpers are plain torch so the host logic is exercised on CPU
with the gloo backend (tests/test_dist_gloo.py).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from . import aggregate_cuda as ac

_REDUCE = {0: "SUM", 1: "MAX", 2: "MIN", 3: "PRODUCT"}
_TORCH_DT = {_abi.DT["float64"]: "float64", _abi.DT["int64"]: "int64", _abi.DT["uint8"]: "uint8"}


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
    """All-reduce each (tensor, reduce) pair in place; reduce in {'sum','max','min','prod'}."""
    import torch.distributed as dist
    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "prod": dist.ReduceOp.PRODUCT}
    for t, red in tables:
        if t.numel():
            dist.all_reduce(t, op=ops[red], group=group)


def _partial_views(lib, req, ws):
    """torch views of the partial tables the library left in the workspace tensor ``ws``."""
    import torch
    parts = (_abi.Partial * 5)()
    k = lib.agc_partials(C.byref(req), parts, 5)
    if k < 0:
        raise RuntimeError(_abi.last_error())
    out = []
    for i in range(k):
        p = parts[i]
        dt = getattr(torch, _TORCH_DT[p.dtype])
        nbytes = p.count * torch.empty((), dtype=dt).element_size()
        out.append((ws[p.offset_bytes:p.offset_bytes + nbytes].view(dt), _abi.REDUCE_NAMES[p.reduce]))
    return out


def aggregate_sharded(group_idx, a, func="sum", size=None, fill_value=0, dtype=None, group=None, **kwargs):
    """``aggregate`` over inputs sharded along N across the ranks of ``group``.

    ``group_idx`` / ``a`` are this rank's contiguous shard (CUDA tensors, Form 1).  ``size`` is required
    (a global maximum would need an extra collective).  Returns the full result on every rank.
    """
    import torch
    import torch.distributed as dist
    if size is None:
        raise ValueError("aggregate_sharded needs an explicit size")
    kwargs = dict(kwargs)
    check = kwargs.get("check", True)
    c = ac._prepare(group_idx, a, func, size, fill_value, "C", dtype, None, kwargs, ac._resolve_size_device)
    if c.kind != "device" or c.plan.form != "flat" or c.base in ac._N_TO_N:
        raise NotImplementedError("aggregate_sharded supports Form 1 reductions; N->N functions are replicas-only")
    lib = _abi.load()
    device = ac._device_of(c)
    lab = ac._labels_on_device(c, device)
    vals = ac._values_on_device(c, device)
    plan = c.plan
    # global element indices only matter for first / last / arg*: everything else skips the all_gather
    needs_offset = c.base in ("first", "last", "argmax", "argmin")
    offset = shard_offsets(plan.n, group)[0] if needs_offset else 0

    compute_out = np.dtype(ac._COMPUTE_AS.get(c.out_dtype.name, c.out_dtype))
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
    req.ddof = float(c.ddof)
    req.index_base = offset
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

        run(base_flags | _abi.F_ACC_ONLY)
        req.flags = base_flags
        merge_tables(_partial_views(lib, req, ws), group)
        if c.base in ("var", "std", "argmax", "argmin"):
            run(base_flags | _abi.F_ACC_ONLY | _abi.F_PASS2 | _abi.F_NO_INIT)
            req.flags = base_flags | _abi.F_PASS2
            merge_tables(_partial_views(lib, req, ws), group)
        if owner_merge:
            # each rank can only gather values it owns: finalise with fill 0, add across ranks, then fill
            idx_table = _partial_views(lib, req, ws)[0][0]
            empty = (idx_table == (2 ** 63 - 1)) | (idx_table < 0)
            f64, i64 = req.fill_f64, req.fill_i64
            req.fill_f64, req.fill_i64 = 0.0, 0
            run(base_flags | _abi.F_FIN_ONLY)
            dist.all_reduce(out.t, op=dist.ReduceOp.SUM, group=group)
            fill_t = torch.tensor(f64 if np.issubdtype(compute_out, np.floating) else i64, dtype=out.t.dtype, device=device)
            out.t = torch.where(empty, fill_t, out.t)
        else:
            run(base_flags | _abi.F_FIN_ONLY)
        if check:
            dist.all_reduce(status, op=dist.ReduceOp.MAX, group=group)
            if int(status.cpu()[_abi.ST_BADLABEL]):           # the only host synchronisation of the call
                raise ValueError("group_idx contains negative indices or indices too large for size")
    if c.out_dtype != compute_out:
        out = ac._Dev(out.t.to(getattr(torch, c.out_dtype.name)), c.out_dtype)
    return ac._as_user_tensor(out)
