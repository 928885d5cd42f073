"""``aggregate_cuda`` -- the B200 implementation of numpy-groupies' ``aggregate``.

A sixth implementation module next to the reference's aggregate_numpy / aggregate_numba /
aggregate_pandas ...: same signature, same Forms 1-5, same dtype / fill_value / exception
rules (``_semantics.py`` restates utils.py), but the per-function kernel layer
(aggregate_numpy.py:20-274) is a set of hand-written sm_100a CUDA kernels reached through
the C ABI of ``csrc/libaggcuda.so`` (include/aggcuda.h).  PyTorch is used for device
buffers and streams only.  There is no CPU fallback: without the shared library or without
a CUDA device every call raises.

Inputs may be numpy arrays / sequences / scalars (copied to the device, result returned as a
numpy array -- what the reference's test-suite does) or CUDA ``torch.Tensor``s (result is a
CUDA tensor; nothing crosses PCIe except the 32-byte status record).

Extra keyword arguments beyond the reference's (it already takes ``**kwargs``):
  check=True          read the device status record after the call (one stream sync) and raise
                      ValueError for bad labels, as the reference does.  Whether labels are valid depends on the labels and
                      ``size`` alone, so a CUDA label tensor that passed is remembered (same tensor object, same torch
                      version counter -- any in-place write bumps it): later calls on it stay asynchronous.
                      ``check="always"`` reads the record back on every call; ``check=False`` never does (bad labels
                      are then ignored silently).
  assume_sorted=False promise non-decreasing flat labels (N->N functions skip the sort; verified).
  fast=True           allow the shared-memory fast kernels (False forces the general path).
  deterministic=False fixed-order mode: floating sum / mean / var / std / sumofsquares / prod add every group's elements in
                      their original order (device radix sort by group, one warp per group, fixed butterfly) -- the same
                      bits on every run; float scans keep the three-kernel variant.  Slower (it sorts); integer, min / max,
                      arg*, first / last and count results never depend on the order.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _abi, _semantics as sem

__all__ = ["aggregate", "unpack", "uaggregate", "aggregate_multi", "group_csr", "step_count", "step_indices",
           "label_contiguous_1d", "relabel_groups_masked", "relabel_groups_unique", "multi_arange", "is_duck_array"]

_torch = None


_LAST = {"regime": 0, "probe": (0, 0, 0)}


def last_regime():
    """Which accumulate kernel the last checked call used (status[3] of the C ABI): 0 general global-table
    kernel, 10+mode k_fast (mode: 0 sum, 1 sum+count, 2 max, 3 min, 4 len, 5 one-word sum), 16 k_len16, 17 / 18
    k_maxfilter max / min, 20+mode k_cache, 40+mode k_form3, 50+mode k_runs, 30 k_fast fused min + max, 60 k_accumulate_small, 70 / 71 k_team (sums /
    sums + counts on a cluster of SMs)."""
    return _LAST["regime"]


def tuning(large=-1, cache_ctas=-1, sum1=-1, smem1=-1, smem2=-1, uncov_pct=-1, scan_chain=-1, team=-1, team_maxu=-1, part_grid=-1):
    """Kernel-selection knobs of libaggcuda (agc_tuning, include/aggcuda.h): for tests and A/B measurements."""
    lib = _abi.load()
    return tuple(lib.agc_tuning(k, int(v)) for k, v in enumerate((large, cache_ctas, sum1, smem1, smem2, uncov_pct, scan_chain, team, team_maxu,
                                                                  part_grid)))


def _t():
    global _torch
    if _torch is None:
        import torch
        _torch = torch
    return _torch


_cuda_ok = False


def _require_cuda():
    global _cuda_ok
    torch = _t()
    if not _cuda_ok:                                        # is_available() costs several microseconds per call: ask once
        if not torch.cuda.is_available():
            raise RuntimeError("aggregate_cuda needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        _cuda_ok = True
    return torch


_NAME = {}


def _dtname(dt):
    """np.dtype(dt).name through a cache (the property assembles the string on every access)."""
    try:
        return _NAME[dt]
    except (KeyError, TypeError):
        name = np.dtype(dt).name
        try:
            _NAME[dt] = name
        except TypeError:
            pass
        return name


# numpy dtype name -> torch dtype used to *carry* the bytes (unsigned types travel as same-width signed)
_CARRIER = {"bool": "bool", "int8": "int8", "uint8": "uint8", "int16": "int16", "uint16": "int16",
            "int32": "int32", "uint32": "int32", "int64": "int64", "uint64": "int64",
            "float32": "float32", "float64": "float64"}
_SIGNED_VIEW = {"uint16": np.int16, "uint32": np.int32, "uint64": np.int64}


class _Dev:
    """A device buffer plus the numpy dtype its bytes mean."""
    __slots__ = ("t", "dtype")

    def __init__(self, t, dtype):
        self.t, self.dtype = t, np.dtype(dtype)

    @property
    def ptr(self):
        return self.t.data_ptr() if self.t.numel() else 0


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


# ---- labels that already passed the bad-label / sortedness check --------------------------------------------------
# The verdict of the check is a function of (labels, size, form) only.  Remembering it per label TENSOR -- weak reference to
# the very object plus its version counter, which every in-place torch operation on it or on a view of it bumps -- lets
# repeated calls on the same labels (group-by over many value columns, the steps of a training loop) run without the one
# host synchronisation a call otherwise has.  A new tensor object, even at the same address, is always checked again.
_VALID = {}


def _valid_sig(c):
    plan = c.plan
    return (plan.form, plan.flat_size, tuple(getattr(plan, "dims", None) or ()), getattr(plan, "axis", None), bool(c.assume_sorted))


def _want_check(c):
    """Does this call have to read the status record back?"""
    if c.check is False:
        return False
    g = c.group_idx
    if c.check == "always" or not (_is_tensor(g) and getattr(g, "is_cuda", False)):
        return True
    e = _VALID.get(id(g))
    if e is None:
        return True
    ref, ver, sigs = e
    if ref() is not g or ver != g._version:
        _VALID.pop(id(g), None)
        return True
    return _valid_sig(c) not in sigs


def _mark_validated(c):
    g = c.group_idx
    if not (_is_tensor(g) and getattr(g, "is_cuda", False)) or c.check is False:
        return
    import weakref
    e = _VALID.get(id(g))
    if e is not None and e[0]() is g and e[1] == g._version:
        e[2].add(_valid_sig(c))
        return
    if len(_VALID) > 256:
        _VALID.clear()
    try:
        _VALID[id(g)] = (weakref.ref(g), g._version, {_valid_sig(c)})
    except TypeError:                                       # (an object that cannot be weakly referenced: never cached)
        pass


_TORCH_NP = {}


def _tensor_np_dtype(t):
    try:
        return _TORCH_NP[t.dtype]
    except KeyError:
        dt = np.dtype(str(t.dtype).split(".")[-1])
        _TORCH_NP[t.dtype] = dt
        return dt


_STAGE = {"bufs": None, "pool": None, "events": None}
_STAGE_CHUNK = 32 << 20          # bytes per pinned staging buffer
_STAGE_BUFS = 3
# host threads that fill the staging buffers (memcpy / label narrowing release the GIL): all cores but two, at most 14
# (measured on the 16-core B200 host, N = 2^28, pageable int64 labels + f32 values: 8 threads 72-77 ms, 12 threads 70-71,
#  14 threads 67-69 -- profiles/r02j_stage_threads.jsonl; two cores stay free for the caller and the driver's copy thread)
# Under torchrun every rank of the node has its own pool: the cores are shared out (LOCAL_WORLD_SIZE).
_STAGE_THREADS = int(os.environ.get("AGC_STAGE_THREADS", 0)) or max(
    2, min(14, (os.cpu_count() or 4) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1)) - 2))
_NARROW_LABELS = os.environ.get("AGC_NARROW_LABELS", "1") != "0"
_NARROW_PINNED = os.environ.get("AGC_NARROW_PINNED", "0") != "0"      # page-locked labels too (measured slower than their plain DMA)
# arrays of at least this many bytes go through the staging ring / as one asynchronous DMA; smaller ones as a plain copy
_STAGE_MIN = int(os.environ.get("AGC_STAGE_MIN", 8 << 20))


def _stage_setup():
    torch = _t()
    if _STAGE["bufs"] is None:
        from concurrent.futures import ThreadPoolExecutor
        _STAGE["bufs"] = [torch.empty(_STAGE_CHUNK, dtype=torch.uint8, pin_memory=True) for _ in range(_STAGE_BUFS)]
        _STAGE["events"] = [None] * _STAGE_BUFS
        _STAGE["pool"] = ThreadPoolExecutor(max_workers=_STAGE_THREADS)
    return _STAGE


def _stage_ring(n_items, item_bytes, fill, dst_u8):
    """Host data -> device through the ring of pinned staging buffers: `fill(stage_u8, lo, hi)` writes items [lo, hi) of
    the source into a staging buffer (called from the pool's threads on slices of a chunk) while the previous chunk's
    asynchronous H2D copy is in flight."""
    torch = _t()
    st = _stage_setup()
    bufs, events, pool = st["bufs"], st["events"], st["pool"]
    per_chunk = _STAGE_CHUNK // item_bytes
    k = 0
    for off in range(0, n_items, per_chunk):
        m = min(per_chunk, n_items - off)
        i = k % _STAGE_BUFS
        if events[i] is not None:
            events[i].synchronize()                        # the copy that last read this buffer is done
        stage = bufs[i].numpy()
        step = -(-m // _STAGE_THREADS)
        step = -(-step // 64) * 64                         # slices start on cache-line multiples of the staging buffer
        futs = [pool.submit(fill, stage[j * item_bytes:min(j + step, m) * item_bytes], off + j, off + min(j + step, m))
                for j in range(0, m, step)]
        for f in futs:
            f.result()
        dst_u8[off * item_bytes:(off + m) * item_bytes].copy_(bufs[i][:m * item_bytes], non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        events[i] = ev
        k += 1


def _upload_staged(src_u8, dst_u8):
    """Pageable host bytes -> device at close to PCIe rate.  (A plain copy from pageable memory runs at ~14 GB/s: the
    driver stages it through one small buffer, single-threaded.)"""
    lib = _abi.load()
    base = src_u8.ctypes.data
    _stage_ring(src_u8.shape[0], 1, lambda stage, lo, hi: lib.agc_host_copy_stream(base + lo, stage.ctypes.data, hi - lo), dst_u8)


def _is_pinned_np(arr):
    """Does this numpy array live in page-locked memory (e.g. a view of a pinned torch tensor)?"""
    import warnings
    try:
        view = arr.view(_SIGNED_VIEW[arr.dtype.name]) if arr.dtype.name in _SIGNED_VIEW else arr
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                 # (read-only memory: we never write through this tensor)
            return bool(_t().from_numpy(view).is_pinned())
    except Exception:                                       # noqa: BLE001
        return False


def _upload_labels_narrowed(arr, device):
    """Flat int64 / uint64 labels with a known size < 2^31 cross PCIe as int32: 8 of the 12 bytes per element of a
    Form 1 call are labels, and the link, not the kernel, bounds a host-array call.  The pool's threads narrow slices
    into the pinned ring (agc_host_narrow_labels: labels outside [0, 2^31) become -1 and are reported as bad labels like
    the originals) while the previous chunk is on the wire; the kernels take int32 labels as they are."""
    torch = _t()
    lib = _abi.load()
    arr = np.ascontiguousarray(arr).reshape(-1)
    n = arr.shape[0]
    code = _abi.dtype_code(arr.dtype)
    base = arr.ctypes.data

    def fill(stage, lo, hi):
        rc = lib.agc_host_narrow_labels(base + lo * 8, code, hi - lo, stage.ctypes.data)
        if rc:
            raise RuntimeError(_abi.last_error())
    with torch.cuda.device(device):
        out = torch.empty(n, dtype=torch.int32, device=device)
        _stage_ring(n, 4, fill, out.view(torch.uint8))
    return _Dev(out, np.int32)


def _upload(arr, device):
    """numpy array -> contiguous device buffer (H2D copy).  Page-locked arrays (e.g. ``torch.empty(..., pin_memory=True)
    .numpy()``) go as one asynchronous DMA; large pageable arrays through the pinned staging ring; small ones directly."""
    torch = _t()
    arr = np.ascontiguousarray(arr)
    dt = arr.dtype
    view = arr.view(_SIGNED_VIEW[dt.name]) if dt.name in _SIGNED_VIEW else arr
    if not view.flags.writeable:                            # torch.from_numpy warns on read-only memory; we never write
        view = view.view()
        try:
            view.flags.writeable = True
        except ValueError:
            view = np.array(view)
    t = torch.from_numpy(view)
    if view.nbytes >= _STAGE_MIN:
        if t.is_pinned():
            return _Dev(t.to(device, non_blocking=True), dt)
        with torch.cuda.device(device):
            out = torch.empty(t.shape, dtype=t.dtype, device=device)
            _upload_staged(view.reshape(-1).view(np.uint8), out.reshape(-1).view(torch.uint8))
        return _Dev(out, dt)
    return _Dev(t.to(device, non_blocking=False), dt)


def _from_tensor(t):
    dt = _tensor_np_dtype(t)
    return _Dev(t.contiguous(), dt)


def _empty(n, dtype, device):
    torch = _t()
    return _Dev(torch.empty(int(n), dtype=getattr(torch, _CARRIER[_dtname(dtype)]), device=device), dtype)


def _convert_f16(dev, to_half):
    """float16 <-> float32 on the device with the library's own kernel (agc_convert_f16): float16 values are computed as
    float32 and float16 results are rounded back once, at the end (check_dtype keeps float16 for float16 input,
    utils.py:435-470).  torch only allocates the destination."""
    torch = _t()
    src = dev.t.contiguous()
    dst = torch.empty(src.shape, dtype=torch.float16 if to_half else torch.float32, device=src.device)
    with torch.cuda.device(src.device):
        rc = _abi.load().agc_convert_f16(src.data_ptr() if src.numel() else 0, dst.data_ptr() if dst.numel() else 0,
                                         src.numel(), 1 if to_half else 0, _stream_ptr(src.device))
    if rc:
        raise RuntimeError(f"agc_convert_f16 failed ({rc}): {_abi.last_error()}")
    return _Dev(dst, np.float16 if to_half else np.float32)


def _download(dev, shape=None):
    arr = dev.t.cpu().numpy()
    if dev.dtype.name in _SIGNED_VIEW:
        arr = arr.view(dev.dtype)
    return arr if shape is None else arr.reshape(shape)


def _as_user_tensor(dev):
    torch = _t()
    nm = _dtname(dev.dtype)
    if nm in _SIGNED_VIEW and hasattr(torch, nm):
        return dev.t.view(getattr(torch, nm))
    return dev.t


# ------------------------------------------------------------------------------------------------
# request assembly
# ------------------------------------------------------------------------------------------------
_DEVICE_OPS = {"sum", "prod", "len", "last", "first", "all", "any", "min", "max", "argmax", "argmin", "mean",
               "sumofsquares", "var", "std", "allnan", "anynan", "cumsum", "cumprod", "cummax", "cummin", "sort"}
_N_TO_N = {"cumsum", "cumprod", "cummax", "cummin", "sort"}
_ORDERED_OPS = {"sum", "sumofsquares", "mean", "var", "std", "prod"}     # what deterministic=True re-routes (floating a)
_COMPUTE_AS = {"float16": np.float32}          # dtypes the kernels do not load natively


class _Call:
    """Everything decided on the host before any kernel runs."""
    pass


def _np_dtype_of_scalar(a):
    return np.dtype(type(a))


def _fill_numbers(fill_value, out_dtype):
    """(double, int64, is_nan) views of fill_value for the finalise kernel."""
    try:
        f = float(fill_value)
    except (TypeError, ValueError):
        f = 0.0
    nan = isinstance(fill_value, (float, np.floating)) and math.isnan(fill_value)
    if np.issubdtype(out_dtype, np.floating):
        i = 0
    else:
        try:
            i = int(np.asarray(fill_value).astype(out_dtype if out_dtype != np.dtype(bool) else np.int64))
            if out_dtype == np.dtype(bool):
                i = int(bool(fill_value))
        except (ValueError, OverflowError):
            i = 0
        if i >= 2 ** 63:
            i -= 2 ** 64
    return f, i, nan


def _prepare(group_idx, a, func, size, fill_value, order, dtype, axis, kwargs, resolve_size):
    """Host-only planning: validates like the reference and decides op / dtypes / index form.
    ``resolve_size(plan, labels)`` supplies the label maxima when ``size`` is None."""
    c = _Call()
    c.check = kwargs.pop("check", True)
    c.assume_sorted = bool(kwargs.pop("assume_sorted", False))
    c.fast = bool(kwargs.pop("fast", True))
    c.deterministic = bool(kwargs.pop("deterministic", False))

    # --- what kind of inputs ------------------------------------------------------------------
    c.device_mode = _is_tensor(group_idx) or _is_tensor(a)
    c.return_cpu = False
    if c.device_mode:
        torch = _t()
        cpu_in = [x for x in (group_idx, a) if _is_tensor(x) and not x.is_cuda]
        if cpu_in:                                          # host torch tensors (e.g. pinned): H2D here, result goes back
            if any(_is_tensor(x) and x.is_cuda for x in (group_idx, a)):
                raise ValueError("group_idx and a must live on the same device")
            _require_cuda()
            dev = torch.device("cuda", torch.cuda.current_device())
            if _is_tensor(group_idx):
                group_idx = group_idx.to(dev, non_blocking=True)
            if _is_tensor(a):
                a = a.to(dev, non_blocking=True)
            c.return_cpu = True
        if not _is_tensor(group_idx):
            group_idx = torch.as_tensor(np.asanyarray(group_idx), device=a.device)
        g_shape, g_dtype = tuple(group_idx.shape), _tensor_np_dtype(group_idx)
        if _is_tensor(a):
            a_scalar, a_shape, a_dtype = (a.ndim == 0), tuple(a.shape), _tensor_np_dtype(a)
            if a_scalar:
                a = a.item()
                a_dtype = _np_dtype_of_scalar(a)
        else:
            a_scalar = isinstance(a, (int, float, complex)) or np.ndim(a) == 0
            if not a_scalar:
                a = torch.as_tensor(np.asanyarray(a), device=group_idx.device)
                a_shape, a_dtype = tuple(a.shape), _tensor_np_dtype(a)
            else:
                a_shape, a_dtype = (), _np_dtype_of_scalar(a if isinstance(a, (int, float, complex)) else a.item())
    else:
        if not isinstance(a, (int, float, complex)):
            a = np.asanyarray(a)
        group_idx = np.asanyarray(group_idx)
        g_shape, g_dtype = group_idx.shape, group_idx.dtype
        a_scalar = isinstance(a, (int, float, complex)) or a.ndim == 0
        a_shape = () if a_scalar else a.shape
        if isinstance(a, (int, float, complex)):
            a_dtype = _np_dtype_of_scalar(a)
        else:
            a_dtype = a.dtype
            if a_scalar:                                   # 0-d array: behave like the python scalar
                a = a.item()
    c.group_idx, c.a, c.a_scalar, c.a_dtype, c.a_shape = group_idx, a, a_scalar, np.dtype(a_dtype), a_shape

    # --- Forms ---------------------------------------------------------------------------------
    plan = sem.plan_index(g_shape, np.issubdtype(g_dtype, np.integer), a_shape, a_scalar, size, order, axis, func)
    if plan.needs_size:
        resolve_size(plan, c)
    c.plan, c.order = plan, order

    # --- function ------------------------------------------------------------------------------
    name = sem.canonical_func(func)
    c.name = name
    if not isinstance(name, str):
        c.kind = "callable"
        c.user_dtype = dtype
        c.fill_value = fill_value
        return c
    nan_form = name.startswith("nan")
    base = name[3:] if nan_form else name
    c.base, c.nan_form = base, nan_form
    if nan_form and a_scalar:
        raise ValueError("nan-version not supported for scalar input.")
    if np.issubdtype(c.a_dtype, np.complexfloating):
        raise NotImplementedError("complex input is not supported by aggregate_cuda")

    n_for_dtype = plan.flat_size if base not in ("cumprod", "cummax", "cummin") else plan.n
    user_dtype = None if dtype is None else np.dtype(dtype)
    dtype_k = sem.result_dtype(user_dtype, name, c.a_dtype, a_scalar, n_for_dtype)
    sem.validate_fill(fill_value, dtype_k, name)
    c.fill_value = fill_value

    if base == "array":
        if kwargs:
            raise TypeError(f"unexpected keyword arguments {sorted(kwargs)} for func 'array'")
        if fill_value is not None and not (np.isscalar(fill_value) or len(fill_value) == 0):
            raise ValueError("fill_value must be None, a scalar or an empty sequence")
        c.kind = "array"
        return c

    # keyword arguments the reference's kernels accept
    c.ddof, c.reverse = 0, False
    if base in ("var", "std"):
        c.ddof = kwargs.pop("ddof", 0)
    elif base == "sort":
        c.reverse = bool(kwargs.pop("reverse", False))
    if kwargs:
        raise TypeError(f"{name}() got unexpected keyword arguments {sorted(kwargs)}")

    if base in ("mean", "var", "std") and a_scalar:
        raise ValueError("cannot take mean / variance with scalar a")
    out_dtype = sem.kernel_out_dtype(name, dtype_k, c.a_dtype, fill_value)
    flags = _abi.F_NANSKIP if nan_form else 0
    if base in ("min", "max"):
        sem.extreme_identity_check(name, fill_value, dtype_k)
        sem.probe_fill(fill_value, out_dtype)
    elif base in ("argmin", "argmax"):
        if not nan_form:                                   # nanarg* first maps a to float (np.where(isnan..))
            sem.extreme_identity_check(name, np.nan, c.a_dtype)
        sem.probe_fill(fill_value, out_dtype)
    elif base in ("first", "last", "prod"):
        sem.probe_fill(fill_value, out_dtype)
    if base in ("sum", "len", "sumofsquares") and fill_value != 0:
        flags |= _abi.F_NEED_TOUCHED
    if base == "prod" and fill_value != 1:
        flags |= _abi.F_NEED_TOUCHED
    if base == "std":
        flags |= _abi.F_SQRT
    if base == "sort" and c.reverse:
        flags |= _abi.F_REVERSE
    if base in ("cumsum",):
        # the reference's in-place `run -= ...; run += a[...]` (aggregate_numpy.py:262-265) raises numpy's
        # TypeError for bool arithmetic and UFuncTypeError (a TypeError) for non same-kind casts
        if dtype_k == np.dtype(bool):
            raise TypeError("numpy boolean subtract, the `-` operator, is not supported (cumsum of bool)")
        mixed = np.result_type(dtype_k, c.a_dtype)
        if not np.can_cast(mixed, dtype_k, "same_kind"):
            raise TypeError(f"Cannot cast ufunc 'add' output from {mixed} to {dtype_k} with casting rule 'same_kind'")
    if isinstance(fill_value, (float, np.floating)) and math.isnan(fill_value):
        flags |= _abi.F_FILL_NAN
    if c.assume_sorted:
        flags |= _abi.F_ASSUME_SORTED
    if not c.fast:
        flags |= _abi.F_NO_FAST
    if c.deterministic:
        flags |= _abi.F_DETERMINISTIC
    if out_dtype == np.dtype(object) or _dtname(out_dtype) not in _abi.DT and _dtname(out_dtype) != "float16":
        raise NotImplementedError(f"output dtype {out_dtype} is not supported by aggregate_cuda")
    if base in _N_TO_N and plan.ndim_idx > 1 and int(np.prod(plan.out_shape, dtype=np.int64)) != plan.n:
        # N->N result cannot be reshaped to the group grid (reference: reshape error, aggregate_numpy.py:366)
        raise ValueError(f"cannot reshape array of size {plan.n} into shape {tuple(plan.out_shape)}")
    c.kind = "device"
    c.flags, c.out_dtype, c.dtype_k = flags, out_dtype, dtype_k
    return c


# ------------------------------------------------------------------------------------------------
# execution
# ------------------------------------------------------------------------------------------------
def _device_of(c):
    torch = _require_cuda()
    for x in (c.group_idx, c.a):
        if _is_tensor(x):
            if not x.is_cuda:
                raise ValueError("torch inputs must live on a CUDA device")
            return x.device
    return torch.device("cuda", torch.cuda.current_device())


def _stream_ptr(device):
    torch = _t()
    try:                                                    # raw handle, ~10x cheaper than building a Stream object
        return torch._C._cuda_getCurrentRawStream(device.index if device.index is not None else torch.cuda.current_device())
    except Exception:                                       # noqa: BLE001
        return torch.cuda.current_stream(device).cuda_stream


def _labels_on_device(c, device):
    if getattr(c, "_labels_dev", None) is None:
        g = c.group_idx
        if _is_tensor(g):
            c._labels_dev = _from_tensor(g)
        else:
            plan = getattr(c, "plan", None)                 # (not there yet while size=None is being resolved)
            g = np.asanyarray(g)
            # pageable int64 labels of a Form 1 call with a known size are narrowed to int32 by the staging threads (they
            # touch every byte anyway); page-locked arrays go as they are, one DMA (narrowing them measured slower: the
            # host threads move 37 GB/s here, the link 54)
            if (_NARROW_LABELS and plan is not None and plan.form == "flat" and not plan.needs_size and 0 < plan.flat_size < 2 ** 31
                    and g.dtype.kind in "iu" and g.dtype.itemsize == 8 and g.nbytes >= _STAGE_MIN and (_NARROW_PINNED or not _is_pinned_np(g))):
                c._labels_dev = _upload_labels_narrowed(g, device)
            else:
                c._labels_dev = _upload(g, device)
    return c._labels_dev


def _resolve_size_device(plan, c):
    """size=None: measure the label maxima with agc_label_minmax (utils.py:515-526 does np.max)."""
    torch = _require_cuda()
    lib = _abi.load()
    device = _device_of(c)
    lab = _labels_on_device(c, device)
    rows = plan.extra.get("rows", 1) if plan.form == "multi" else 1
    per_row = lab.t.numel() // max(rows, 1)
    if per_row == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    mm = torch.empty((rows, 2), dtype=torch.int64, device=device)
    es = lab.dtype.itemsize
    with torch.cuda.device(device):
        for r in range(rows):
            rc = lib.agc_label_minmax(lab.ptr + r * per_row * es, _abi.dtype_code(lab.dtype), per_row,
                                      mm.data_ptr() + r * 16, _stream_ptr(device))
            if rc:
                raise RuntimeError(_abi.last_error())
    mm = mm.cpu().numpy()
    if (mm[:, 0] < 0).any():
        raise ValueError("negative indices not supported")
    if plan.form == "flat":
        plan.flat_size = int(mm[0, 1]) + 1
    elif plan.form == "axis":
        sem.finish_axis_plan(plan, int(mm[0, 1]) + 1)
    else:
        sem.finish_multi_plan(plan, [int(v) + 1 for v in mm[:, 1]])


def _build_index(c, lab):
    plan = c.plan
    ix = _abi.Index()
    ix.form = {"flat": _abi.FORM_FLAT, "axis": _abi.FORM_AXIS, "multi": _abi.FORM_MULTI}[plan.form]
    ix.dtype = _abi.dtype_code(lab.dtype)
    ix.labels = lab.ptr
    if plan.form != "flat":
        nd = len(plan.dims)
        if nd > _abi.MAX_NDIM:
            raise NotImplementedError(f"more than {_abi.MAX_NDIM} dimensions are not supported")
        ix.ndim, ix.axis, ix.axis_size = nd, plan.axis, plan.axis_size
        for d in range(nd):
            ix.dims[d], ix.strides[d] = plan.dims[d], plan.strides[d]
    return ix


def _values_on_device(c, device):
    if c.a_scalar:
        return None
    a = c.a
    if _is_tensor(a):
        dev = _from_tensor(a)
        if _dtname(dev.dtype) in _COMPUTE_AS:
            dev = _convert_f16(dev, False)
        dev.t = dev.t.reshape(-1)
        return dev
    arr = np.ascontiguousarray(a)
    if arr.dtype.name in _COMPUTE_AS:
        arr = arr.astype(_COMPUTE_AS[arr.dtype.name])
    return _upload(arr.reshape(-1), device)


def _run_device(c):
    torch = _require_cuda()
    lib = _abi.load()
    device = _device_of(c)
    plan = c.plan
    vals = _values_on_device(c, device)                    # first: a page-locked `a` is one asynchronous DMA, already on the
    lab = _labels_on_device(c, device)                      # wire while the host threads narrow / stage the labels
    base = c.base
    n_out = plan.n if base in _N_TO_N else plan.flat_size
    out_np_dtype = c.out_dtype
    compute_out = np.dtype(_COMPUTE_AS.get(_dtname(out_np_dtype), out_np_dtype))

    req = _abi.Request()
    req.op = _abi.OP[base]
    req.flags = c.flags
    a_dt = np.dtype(_COMPUTE_AS.get(_dtname(c.a_dtype), c.a_dtype))
    req.a_dtype = _abi.dtype_code(a_dt)
    req.out_dtype = _abi.dtype_code(compute_out)
    if vals is None:
        req.a = None
        req.a_scalar_f64 = float(c.a)
        req.a_scalar_i64 = int(c.a) if not isinstance(c.a, (float, np.floating)) else 0
    else:
        req.a = vals.ptr
    req.index = _build_index(c, lab)
    req.n, req.size = plan.n, plan.flat_size
    f64, i64, _ = _fill_numbers(c.fill_value, compute_out)
    req.fill_f64, req.fill_i64 = f64, i64
    req.ddof = float(c.ddof)
    req.index_base = 0
    req.arg_stride, req.arg_len = plan.arg_stride, plan.arg_len

    with torch.cuda.device(device):
        out = _empty(n_out, compute_out, device)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        req.out = out.ptr
        req.status = status.data_ptr()
        ws_bytes = lib.agc_workspace_bytes(C.byref(req))
        if ws_bytes < 0:
            raise RuntimeError(_abi.last_error())
        ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
        req.workspace, req.workspace_bytes = ws.data_ptr(), ws_bytes
        ordered = (c.deterministic and base in _ORDERED_OPS and vals is not None and np.issubdtype(a_dt, np.floating) and plan.n > 0)
        if ordered:
            # fixed-order mode: group the element positions (stable radix sort), then one warp per group adds in element order
            order_t, offsets_t = _group_order(c)
            req.flags = c.flags | _abi.F_ACC_ONLY
            rc = lib.agc_ordered_accumulate(C.byref(req), order_t.data_ptr(), offsets_t.data_ptr(), _stream_ptr(device))
            if rc == 0:
                req.flags = c.flags | _abi.F_FIN_ONLY
                rc = lib.agc_aggregate(C.byref(req), _stream_ptr(device))
        else:
            rc = lib.agc_aggregate(C.byref(req), _stream_ptr(device))
        if rc:
            msg = _abi.last_error()
            if rc == -20:
                raise NotImplementedError(msg)
            raise RuntimeError(f"agc_aggregate failed ({rc}): {msg}")
        if _want_check(c):
            st = status.cpu().numpy()
            if st[_abi.ST_BADLABEL]:
                raise ValueError("group_idx contains negative indices or indices too large for size")
            if st[_abi.ST_UNSORTED]:
                raise ValueError("assume_sorted=True but group_idx is not sorted")
            _mark_validated(c)
            c.regime = int(st[_abi.ST_REGIME])
            _LAST["regime"] = c.regime
            _LAST["probe"] = tuple(int(v) for v in st[4:7])
    if out_np_dtype != compute_out:                         # float16 results
        out = _convert_f16(out, True)
    return out


def _finish_shape_np(c, arr):
    plan = c.plan
    if plan.ndim_idx > 1:
        shape = plan.out_shape
        if int(np.prod(shape, dtype=np.int64)) != arr.size:
            raise ValueError(f"cannot reshape array of size {arr.size} into shape {tuple(shape)}")
        arr = arr.reshape(shape, order=c.order)
    return arr


def _finish_shape_torch(c, t):
    plan = c.plan
    if plan.ndim_idx > 1:
        shape = tuple(plan.out_shape)
        if int(np.prod(shape, dtype=np.int64)) != t.numel():
            raise ValueError(f"cannot reshape tensor of size {t.numel()} into shape {shape}")
        if c.order == "F":
            t = t.reshape(shape[::-1]).permute(*range(len(shape) - 1, -1, -1))
        else:
            t = t.reshape(shape)
    return t


def _group_order(c):
    """Stable device sort of element positions by flat group: (order[n], offsets[size+1]) as numpy."""
    torch = _require_cuda()
    lib = _abi.load()
    device = _device_of(c)
    plan = c.plan
    lab = _labels_on_device(c, device)
    ix = _build_index(c, lab)
    with torch.cuda.device(device):
        order = torch.empty(plan.n, dtype=torch.int64, device=device)
        offsets = torch.empty(plan.flat_size + 1, dtype=torch.int64, device=device)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        ws_bytes = lib.agc_group_order_workspace_bytes(plan.n, plan.flat_size)
        ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
        rc = lib.agc_group_order(C.byref(ix), plan.n, plan.flat_size, order.data_ptr(), offsets.data_ptr(),
                                 ws.data_ptr(), ws_bytes, status.data_ptr(), _stream_ptr(device))
        if rc:
            msg = _abi.last_error()
            if rc == -20:
                raise NotImplementedError(msg)
            raise RuntimeError(f"agc_group_order failed ({rc}): {msg}")
        if status.cpu().numpy()[_abi.ST_BADLABEL]:
            raise ValueError("group_idx contains negative indices or indices too large for size")
    return order, offsets


def _host_values(c):
    if c.a_scalar:
        raise ValueError("scalar a is not supported for this function")
    a = c.a
    arr = a.detach().cpu().numpy() if _is_tensor(a) else np.asarray(a)
    return arr.reshape(-1)


def _run_array(c):
    """func='array' (aggregate_numpy.py:217-227): grouping order from the device radix sort, the
    object array itself can only live on the host."""
    order, offsets = _group_order(c)
    vals = _host_values(c)
    order, offsets = order.cpu().numpy(), offsets.cpu().numpy()
    pieces = np.split(vals[order], offsets[1:-1])
    ret = np.asanyarray(pieces, dtype="object")
    fill = c.fill_value
    if fill is None or np.isscalar(fill):
        counts = np.diff(offsets)
        ret[counts == 0] = fill
    return ret


def _run_callable(c):
    """Arbitrary per-group callables (aggregate_numpy.py:230-241): device sort, host loop."""
    order, offsets = _group_order(c)
    vals = _host_values(c)
    order, offsets = order.cpu().numpy(), offsets.cpu().numpy()
    sorted_vals = vals[order]
    ret = np.full(c.plan.flat_size, c.fill_value, dtype=c.user_dtype or np.float64)
    for g in range(c.plan.flat_size):
        lo, hi = offsets[g], offsets[g + 1]
        if hi > lo:
            ret[g] = c.name(sorted_vals[lo:hi])
    return ret


def aggregate(group_idx, a, func="sum", size=None, fill_value=0, order="C", dtype=None, axis=None, **kwargs):
    """Group-wise reduction / scan of ``a`` by the integer labels ``group_idx`` on a B200.

    Same contract as ``numpy_groupies.aggregate_numpy.aggregate`` (aggregate_numpy.py:370-392;
    parameters documented in utils.py:6-56): Forms 1-5 of ``group_idx``/``a``, ``func`` as string,
    builtin, numpy function or arbitrary callable, ``size``, ``fill_value``, ``order``, ``dtype``,
    ``axis``, plus ``ddof`` (var/std) and ``reverse`` (sort).
    """
    c = _prepare(group_idx, a, func, size, fill_value, order, dtype, axis, dict(kwargs), _resolve_size_device)
    if c.kind == "device":
        out = _run_device(c)
        if c.device_mode:
            res = _finish_shape_torch(c, _as_user_tensor(out))
            return res.cpu() if c.return_cpu else res
        return _finish_shape_np(c, _download(out))
    if c.kind == "array":
        ret = _run_array(c)
    else:
        ret = _run_callable(c)
    return _finish_shape_np(c, ret)


def _unpack_dev(lab, tab):
    """agc_unpack on device buffers: tab[lab] as a new _Dev of tab's dtype."""
    torch = _require_cuda()
    lib = _abi.load()
    device = lab.t.device
    n = lab.t.numel()
    with torch.cuda.device(device):
        out = _empty(n, tab.dtype, device)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        rc = lib.agc_unpack(tab.ptr, tab.dtype.itemsize, tab.t.numel(), lab.ptr, _abi.dtype_code(lab.dtype), n,
                            out.ptr, status.data_ptr(), _stream_ptr(device))
        if rc:
            raise RuntimeError(_abi.last_error())
        if n and status.cpu().numpy()[_abi.ST_BADLABEL]:
            raise IndexError("index out of bounds in unpack")
    return out


def unpack(group_idx, ret):
    """``ret[group_idx]`` on the device (utils.py:548-553): the packed per-group result broadcast back to the elements."""
    torch = _require_cuda()
    device_mode = _is_tensor(group_idx) or _is_tensor(ret)
    if device_mode:
        device = next(x.device for x in (group_idx, ret) if _is_tensor(x) and x.is_cuda)
    else:
        device = torch.device("cuda", torch.cuda.current_device())
    lab = _as_device_ints(group_idx, device)
    tab = _from_tensor(ret if ret.is_cuda else ret.to(device)) if _is_tensor(ret) else _upload(np.asanyarray(ret), device)
    if lab.dtype.kind not in "iu":
        raise IndexError("arrays used as indices must be of integer type")
    out = _unpack_dev(lab, tab)
    shape = tuple(group_idx.shape) if hasattr(group_idx, "shape") else np.shape(group_idx)
    if device_mode:
        return _as_user_tensor(out).reshape(shape)
    return _download(out).reshape(shape)


# ------------------------------------------------------------------------------------------------
# test hook: host-only planning (no device), used by tests/test_semantics.py to compare dtype /
# shape / exception decisions with the reference's recorded answers on a machine without a GPU.
# ------------------------------------------------------------------------------------------------
def _plan_only_for_tests(group_idx, a, func="sum", size=None, fill_value=0, order="C", dtype=None, axis=None, **kwargs):
    def resolve_on_host(plan, c):
        g = np.asanyarray(c.group_idx)
        if g.size == 0:
            raise ValueError("zero-size array to reduction operation maximum which has no identity")
        if (g < 0).any():
            raise ValueError("negative indices not supported")
        if plan.form == "flat":
            plan.flat_size = int(g.max()) + 1
        elif plan.form == "axis":
            sem.finish_axis_plan(plan, int(g.max()) + 1)
        else:
            sem.finish_multi_plan(plan, [int(v) + 1 for v in g.max(axis=1)])
    return _prepare(group_idx, a, func, size, fill_value, order, dtype, axis, dict(kwargs), resolve_on_host)


# ------------------------------------------------------------------------------------------------
# the callers either side of the path: uaggregate, fused multi-statistic aggregation,
# device-side grouping (CSR) for `array` / generic callables, label utilities
# ------------------------------------------------------------------------------------------------
def uaggregate(group_idx, a, **kwargs):
    """``unpack(group_idx, aggregate(group_idx, a, **kwargs))`` (numpy_groupies/__init__.py:42-43): the per-group
    result broadcast back to the elements.  Device tensors stay on the device between the two steps."""
    return unpack(group_idx, aggregate(group_idx, a, **kwargs))


# func -> (family, what the finalise step is asked for).  One accumulate pass per family serves all its members.
_MOMENTS = {"sum": "sum", "len": "len", "mean": "mean", "var": "var", "std": "std"}
_EXTREMES = {"min", "max"}


def aggregate_multi(group_idx, a, funcs, size=None, fill_value=0, order="C", dtype=None, axis=None, **kwargs):
    """Several statistics of the same ``(group_idx, a)`` from as few passes over the data as possible -- the fused
    aggregation the reference's README (:185, "min+max in one pass") names as the missed optimisation.

    ``funcs`` is a sequence of function names; the result is a dict ``{name: array}`` whose entries equal
    ``aggregate(group_idx, a, name, size=..., fill_value=..., ...)`` one by one.  Fused families:
      * sum / len / mean / var / std (and their nan-forms, as a family of their own) on floating ``a``: ONE accumulate
        pass leaves (sum, count) per group, var / std add the reference's second pass (aggregate_numpy.py:180-187);
        every member is just another finalise over the same tables;
      * min + max: one pass that feeds both key tables (AGC_OP_MINMAX).
    Everything else, integer ``a`` for sum (its accumulator is an exact int64) and ``dtype=`` overrides run as separate
    calls.  ``size`` should be given: with ``size=None`` every member would measure the labels again.
    """
    names = [sem.canonical_func(f) for f in funcs]
    if any(not isinstance(nm, str) for nm in names):
        raise TypeError("aggregate_multi takes function names")
    out = {}
    todo = list(dict.fromkeys(names))
    ddof = kwargs.get("ddof", 0)
    plain_kwargs = {k: v for k, v in kwargs.items() if k != "ddof"}

    def single(nm):
        kw = dict(plain_kwargs)
        if nm.replace("nan", "", 1) in ("var", "std"):
            kw["ddof"] = ddof
        return aggregate(group_idx, a, nm, size=size, fill_value=fill_value, order=order, dtype=dtype, axis=axis, **kw)

    a_is_float = (np.issubdtype(_tensor_np_dtype(a), np.floating) if _is_tensor(a)
                  else np.issubdtype(np.asanyarray(a).dtype, np.floating)) if not isinstance(a, (int, float)) else False
    for nan_form in (False, True):
        pre = "nan" if nan_form else ""
        mom = [nm for nm in todo if nm.startswith(pre) and nm[len(pre):] in _MOMENTS and (nan_form or not nm.startswith("nan"))]
        if len(mom) >= 2 and a_is_float and dtype is None and size is not None:
            for nm, res in _fused_family(group_idx, a, mom, "moments", size, fill_value, order, axis, ddof, plain_kwargs).items():
                out[nm] = res
        ext = [nm for nm in todo if nm.startswith(pre) and nm[len(pre):] in _EXTREMES and (nan_form or not nm.startswith("nan"))]
        if len(ext) == 2 and dtype is None and size is not None:
            for nm, res in _fused_family(group_idx, a, ext, "extremes", size, fill_value, order, axis, ddof, plain_kwargs).items():
                out[nm] = res
    for nm in todo:
        if nm not in out:
            out[nm] = single(nm)
    return {nm: out[nm] for nm in names}


def _fused_family(group_idx, a, members, family, size, fill_value, order, axis, ddof, kwargs):
    """One accumulate pass (two for var / std), one finalise per member.  Every member is planned by the same host
    code as a single call (dtype, fill and error rules are the reference's), only the kernels are shared."""
    torch = _require_cuda()
    lib = _abi.load()
    calls = {}
    for nm in members:
        kw = dict(kwargs)
        if nm.replace("nan", "", 1) in ("var", "std"):
            kw["ddof"] = ddof
        calls[nm] = _prepare(group_idx, a, nm, size, fill_value, order, None, axis, kw, _resolve_size_device)
    c0 = calls[members[0]]
    device = _device_of(c0)
    lab = _labels_on_device(c0, device)
    vals = _values_on_device(c0, device)
    plan = c0.plan
    if vals is None:
        raise ValueError("aggregate_multi needs an array a")
    bases = {nm: c.base for nm, c in calls.items()}
    need_pass2 = any(b in ("var", "std") for b in bases.values())
    if family == "moments":
        acc_op = "var" if need_pass2 else "mean"
    else:
        acc_op = "minmax"

    results = {}
    with torch.cuda.device(device):
        status = torch.zeros(8, dtype=torch.int32, device=device)
        stream = _stream_ptr(device)
        req = _abi.Request()
        a_dt = np.dtype(_COMPUTE_AS.get(_dtname(c0.a_dtype), c0.a_dtype))
        req.a_dtype = _abi.dtype_code(a_dt)
        req.a = vals.ptr
        req.index = _build_index(c0, lab)
        req.n, req.size = plan.n, plan.flat_size
        req.index_base = 0
        req.arg_stride, req.arg_len = plan.arg_stride, plan.arg_len
        req.status = status.data_ptr()
        nan_flag = _abi.F_NANSKIP if c0.nan_form else 0

        def call(op, flags, c=None, out=None):
            req.op = _abi.OP[op]
            req.flags = flags
            cc = c or c0
            compute_out = np.dtype(_COMPUTE_AS.get(_dtname(cc.out_dtype), cc.out_dtype))
            req.out_dtype = _abi.dtype_code(compute_out)
            req.fill_f64, req.fill_i64, _ = _fill_numbers(cc.fill_value, compute_out)
            req.ddof = float(getattr(cc, "ddof", 0))
            req.out = out.ptr if out is not None else 0
            rc = lib.agc_aggregate(C.byref(req), stream)
            if rc:
                raise RuntimeError(f"agc_aggregate failed ({rc}): {_abi.last_error()}")

        # workspace: the reduction layout does not depend on the op
        req.op, req.flags, req.out_dtype, req.out = _abi.OP["mean"], _abi.F_ACC_ONLY, _abi.dtype_code(np.dtype(np.float64)), 0
        ws_bytes = lib.agc_workspace_bytes(C.byref(req))
        if ws_bytes < 0:
            raise RuntimeError(_abi.last_error())
        ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
        req.workspace, req.workspace_bytes = ws.data_ptr(), ws_bytes

        base_flags = nan_flag | (0 if c0.fast else _abi.F_NO_FAST)
        call(acc_op, base_flags | _abi.F_ACC_ONLY)                       # pass 1: (sum, count) or both key tables

        def finish(nm):
            c = calls[nm]
            compute_out = np.dtype(_COMPUTE_AS.get(_dtname(c.out_dtype), c.out_dtype))
            o = _empty(plan.flat_size, compute_out, device)
            flags = c.flags | _abi.F_FIN_ONLY
            if family == "moments" and c.base == "sum":
                flags |= _abi.F_TOUCHED_BY_COUNT
            if family == "extremes" and c.base == "min":
                flags |= _abi.F_KEYS_IN_T2
            call(c.base, flags, c, o)
            if c.out_dtype != compute_out:
                o = _convert_f16(o, True)
            return o

        first = [nm for nm in members if bases[nm] not in ("var", "std")]
        for nm in first:                                                 # sum / len / mean read the pass-1 tables
            results[nm] = finish(nm)
        if need_pass2:                                                   # then T0 turns into the means and pass 2 runs
            call("var", base_flags | _abi.F_ACC_ONLY | _abi.F_PASS2 | _abi.F_NO_INIT)
            for nm in members:
                if bases[nm] in ("var", "std"):
                    results[nm] = finish(nm)
        if _want_check(c0):
            st = status.cpu().numpy()
            if st[_abi.ST_BADLABEL]:
                raise ValueError("group_idx contains negative indices or indices too large for size")
            _mark_validated(c0)
            _LAST["regime"] = int(st[_abi.ST_REGIME])
    final = {}
    for nm, o in results.items():
        c = calls[nm]
        if c.device_mode:
            res = _finish_shape_torch(c, _as_user_tensor(o))
            final[nm] = res.cpu() if c.return_cpu else res
        else:
            final[nm] = _finish_shape_np(c, _download(o))
    return final


def group_csr(group_idx, a=None, size=None, order="C", axis=None, check=True):
    """Device-side grouping for ``func='array'`` / arbitrary callables (aggregate_numpy.py:217-241,
    aggregate_numba.py:228-291) WITHOUT the trip through Python objects: a stable radix sort of the element positions
    by flat group.  Returns ``(order, offsets, values)`` as CUDA tensors: ``order[offsets[g]:offsets[g+1]]`` are the
    positions of group ``g`` in their original order, ``values = a.ravel()[order]`` (``None`` when ``a`` is None) --
    i.e. a CSR layout any device-side per-group function can walk."""
    torch = _require_cuda()
    probe = a if a is not None else 0
    c = _prepare(group_idx, probe, "array", size, 0, order, None, axis, {"check": check}, _resolve_size_device)
    order_t, offsets_t = _group_order(c)
    values = None
    if a is not None and not c.a_scalar:
        dev = _values_on_device(c, _device_of(c))
        values = _as_user_tensor(_unpack_dev(_Dev(order_t, np.int64), dev))
    return order_t, offsets_t, values


def _as_device_ints(x, device):
    if _is_tensor(x):
        return _from_tensor(x if x.is_cuda else x.to(device))
    arr = np.asanyarray(x)
    return _upload(arr, device)


def _label_scan(mode, x, out_dtype=np.int64, out_len=None, want_total=False):
    """agc_label_scan on a device buffer ``x`` (a _Dev): returns (out _Dev | None, total int | None)."""
    torch = _require_cuda()
    lib = _abi.load()
    device = x.t.device
    n = x.t.numel()
    with torch.cuda.device(device):
        out = _empty(out_len if out_len is not None else n, out_dtype, device) if mode != _abi.LB_STEP_COUNT else None
        total = torch.zeros(1, dtype=torch.int64, device=device)
        ws_bytes = lib.agc_label_scan_workspace_bytes(n)
        ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=device)
        rc = lib.agc_label_scan(mode, x.ptr, _abi.dtype_code(x.dtype), n, out.ptr if out is not None else 0,
                                _abi.dtype_code(np.dtype(out_dtype)), total.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr(device))
        if rc:
            raise RuntimeError(_abi.last_error())
        tot = int(total.item()) if want_total else None
    return out, tot


def _check_int_1d(x, what):
    nd = x.ndim if hasattr(x, "ndim") else np.ndim(x)
    if nd != 1:
        raise ValueError(f"{what} is supposed to be 1d array.")


def _finish_like(x, dev):
    return _as_user_tensor(dev) if _is_tensor(x) else _download(dev)


def step_count(group_idx):
    """Number of runs of equal labels in ``group_idx`` (aggregate_numba.py:577-588), counted on the device."""
    torch = _require_cuda()
    device = group_idx.device if _is_tensor(group_idx) and group_idx.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = _as_device_ints(group_idx, device)
    if x.t.numel() == 0:
        return 0
    return _label_scan(_abi.LB_STEP_COUNT, x, want_total=True)[1]


def step_indices(group_idx):
    """Edges of the runs of equal labels (aggregate_numba.py:591-605): ``[0, ..., len(group_idx)]``."""
    torch = _require_cuda()
    device = group_idx.device if _is_tensor(group_idx) and group_idx.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = _as_device_ints(group_idx, device)
    n = x.t.numel()
    if n == 0:
        # the reference writes indices[0] = 0 and indices[-1] = size into a one-element array
        return _finish_like(group_idx, _upload(np.zeros(1, dtype=np.int64), device))
    out, total = _label_scan(_abi.LB_STEPS, x, out_len=n + 1, want_total=True)
    out.t = out.t[:total + 1]
    return _finish_like(group_idx, out)


def label_contiguous_1d(X):
    """Label contiguous blocks of a 1-d mask / of equal non-zero values with 1, 2, 3 ... (utils.py:597-632)."""
    torch = _require_cuda()
    _check_int_1d(X, "this")
    device = X.device if _is_tensor(X) and X.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = _as_device_ints(X, device)
    if x.dtype.kind not in "biu":
        raise NotImplementedError("label_contiguous_1d on the device takes bool or integer input")
    out, _ = _label_scan(_abi.LB_CONTIG, x)
    return _finish_like(X, out)


def _relabel(g, keep, like):
    table, _ = _label_scan(_abi.LB_KEEP, keep, out_dtype=g.dtype)          # table[k] = new number of group k (0 if removed)
    return _finish_like(like, _unpack_dev(g, table))


def relabel_groups_masked(group_idx, keep_group):
    """Remove the groups whose ``keep_group`` entry is falsy and renumber the rest without gaps (utils.py:651-681)."""
    torch = _require_cuda()
    device = group_idx.device if _is_tensor(group_idx) and group_idx.is_cuda else torch.device("cuda", torch.cuda.current_device())
    g = _as_device_ints(group_idx, device)
    if _is_tensor(keep_group):
        k = _as_device_ints(keep_group, device)
    else:
        k = _upload(np.asanyarray(keep_group).astype(bool), device)
    return _relabel(g, k, group_idx)


def relabel_groups_unique(group_idx):
    """Renumber the labels that actually occur (plus group 0) without gaps, keeping their order (utils.py:635-648)."""
    torch = _require_cuda()
    lib = _abi.load()
    device = group_idx.device if _is_tensor(group_idx) and group_idx.is_cuda else torch.device("cuda", torch.cuda.current_device())
    g = _as_device_ints(group_idx, device)
    n = g.t.numel()
    if n == 0:
        raise ValueError("zero-size array to reduction operation maximum which has no identity")
    with torch.cuda.device(device):
        mm = torch.empty(2, dtype=torch.int64, device=device)
        rc = lib.agc_label_minmax(g.ptr, _abi.dtype_code(g.dtype), n, mm.data_ptr(), _stream_ptr(device))
        if rc:
            raise RuntimeError(_abi.last_error())
        lo, hi = (int(v) for v in mm.cpu())
        if lo < 0:
            raise IndexError("negative labels in relabel_groups_unique")
        keep = torch.zeros(hi + 1, dtype=torch.uint8, device=device)
        status = torch.zeros(8, dtype=torch.int32, device=device)
        rc = lib.agc_mark_present(g.ptr, _abi.dtype_code(g.dtype), n, hi + 1, keep.data_ptr(), status.data_ptr(), _stream_ptr(device))
        if rc:
            raise RuntimeError(_abi.last_error())
    return _relabel(g, _Dev(keep, np.uint8), group_idx)


def multi_arange(n):
    """``hstack([arange(n_i) for n_i in n])`` (utils.py:572-594) on the device: running sums of ``n``, then every output
    position finds its segment by bisection."""
    torch = _require_cuda()
    lib = _abi.load()
    _check_int_1d(n, "n")
    device = n.device if _is_tensor(n) and n.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = _as_device_ints(n, device)
    m = x.t.numel()
    if m == 0:
        raise IndexError("index -1 is out of bounds for axis 0 with size 0")
    incl, total = _label_scan(_abi.LB_CUMSUM, x, want_total=True)
    with torch.cuda.device(device):
        out = _empty(total, np.int64, device)
        rc = lib.agc_multi_arange(incl.ptr, m, total, out.ptr, _stream_ptr(device))
        if rc:
            raise RuntimeError(_abi.last_error())
    return _finish_like(n, out)


def is_duck_array(value):
    """What the facade accepts as array input (utils.py:684-695): numpy arrays, anything that implements the numpy array
    protocols, and -- for this implementation -- torch tensors (they stay on their device)."""
    if isinstance(value, np.ndarray) or _is_tensor(value):
        return True
    return (hasattr(value, "ndim") and hasattr(value, "shape") and hasattr(value, "dtype")
            and hasattr(value, "__array_function__") and hasattr(value, "__array_ufunc__"))
