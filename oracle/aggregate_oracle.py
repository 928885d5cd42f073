"""CPU oracle for the ``aggregate`` hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This module is a from-scratch numpy restatement of what the reference computes on the
path ``aggregate_numpy.aggregate`` (and, for cumprod/cummax/cummin, what
``aggregate_numba.aggregate`` computes).  It exists only so that ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs can check and time the CUDA path on machines where ``/root/reference`` does not
exist.  Nothing under ``numpy_groupies_b200/`` imports it.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the real reference
(``/root/reference/numpy_groupies``: aggregate_numpy + aggregate_numba) in the build
container and stores its answers in ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``
replays every stored case through this module and demands identical dtype, shape,
values (bit-exact for integers / bools / indices, 1e-12 relative for floats) and
identical exception classes.

Every function cites the reference location it restates (paths relative to
``/root/reference/numpy_groupies``).  The arithmetic deliberately uses the same numpy
primitives as the reference (``np.bincount`` accumulates in float64 in input order,
``ufunc.at`` for min/max/prod, stable sorts for the N->N functions) because those
primitives *are* the reference's numerics.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# 1. Function-name resolution                        (utils.py:58-179, aggregate_numpy.py:277-305)
# --------------------------------------------------------------------------------------
_NO_NAN_FORM = frozenset(("sort", "rsort", "array", "allnan", "anynan"))      # utils.py:59

# canonical name -> every spelling the reference accepts for it (utils.py:62-131)
_SPELLINGS = {
    "any": ("or", any, np.any),
    "all": ("and", all, np.all),
    "sum": ("add", "plus", sum, np.add, np.sum),
    "len": ("count", len),
    "prod": ("multiply", "product", "times", np.multiply, np.prod),
    "max": ("amax", "maximum", max, np.amax, np.max, np.maximum),
    "min": ("amin", "minimum", min, np.amin, np.min, np.minimum),
    "array": ("split", "splice", slice, list, np.array, np.asarray),
    "sort": ("sorted", "asort", "asorted", "rsorted", "dsort", sorted, np.sort),
    "rsort": ("dsorted",),
    "argmax": (np.argmax,), "argmin": (np.argmin,),
    "mean": (np.mean,), "std": (np.std,), "var": (np.var,),
    "cumsum": (np.cumsum,), "cumprod": (np.cumprod,),
    "nansum": (np.nansum,), "nanprod": (np.nanprod,), "nanmean": (np.nanmean,),
    "nanvar": (np.nanvar,), "nanstd": (np.nanstd,), "nanmax": (np.nanmax,),
    "nanmin": (np.nanmin,), "nanargmax": (np.nanargmax,), "nanargmin": (np.nanargmin,),
    "nancumsum": (np.nancumsum,),
    "first": (), "last": (), "allnan": (), "anynan": (), "sumofsquares": (),
    "cummax": (), "cummin": (),
}


def _build_alias_table():
    table = {}
    for canon, spellings in _SPELLINGS.items():
        table[canon] = canon
        for s in spellings:
            table[s] = canon
    for canon in list(_SPELLINGS):                                              # utils.py:151-155
        if canon not in _NO_NAN_FORM and not canon.startswith("nan"):
            table["nan" + canon] = "nan" + canon
    return table


ALIASES = _build_alias_table()

# what aggregate_numpy implements (aggregate_numpy.py:277-305) plus the three N->N
# functions only aggregate_numba has (aggregate_numba.py:491-500, tests/__init__.py:33)
_NUMPY_FUNCS = ("min max sum prod last first all any mean std var anynan allnan sort array "
                "argmax argmin len cumsum sumofsquares").split()
_NUMBA_ONLY = ("cumprod", "cummax", "cummin")
IMPLEMENTED = set(_NUMPY_FUNCS) | set(_NUMBA_ONLY)
IMPLEMENTED |= {"nan" + f for f in IMPLEMENTED if f not in _NO_NAN_FORM}


def resolve_func(func):
    """utils.py:163-179 -- canonical name, or the callable itself."""
    try:
        name = ALIASES[func]
    except KeyError:
        if callable(func):
            return func
        raise ValueError(f"func {func!r} is neither a valid function string nor a callable object")
    if name in IMPLEMENTED:
        return name
    raise NotImplementedError("No such function available")


# --------------------------------------------------------------------------------------
# 2. dtype / fill_value rules                                      (utils.py:182-374)
# --------------------------------------------------------------------------------------
_INT_LADDER = {"bool": np.int8, "uint8": np.int16, "int8": np.int16, "uint16": np.int32,
               "int16": np.int32, "uint32": np.int64, "int32": np.int64}      # utils.py:187-195
_FLOAT_LADDER = {"float16": np.float32, "float32": np.float64, "float64": np.complex64,
                 "complex64": np.complex128}                                   # utils.py:197-202


def _represents(x, dt):
    """utils.py:211-219: can dtype ``dt`` hold ``x`` unchanged (NaN always counts as held)."""
    try:
        with np.errstate(invalid="ignore"):
            conv = np.array(x).astype(dt)
    except (ValueError, OverflowError, RuntimeWarning):
        return False
    return conv == x or np.isnan(x)


def smallest_dtype_for(x, dtype=np.bool_):
    """utils.py:205-238 (``minimum_dtype``)."""
    dt = np.dtype(dtype)
    if _represents(x, dt):
        return dt
    floating = np.issubdtype(dt, np.inexact)
    ladder = _FLOAT_LADDER if floating else _INT_LADDER
    while True:
        nxt = ladder.get(dt.name)
        if nxt is None:
            if floating:
                raise ValueError(f"Can not determine dtype of {x!r}")
            return np.dtype(np.float32)                   # integer ladder exhausted -> float32
        dt = np.dtype(nxt)
        if _represents(x, dt):
            return dt


def _dtype_for_fill(fill_value, dtype, a):
    """utils.py:241-244 (``minimum_dtype_scalar``)."""
    if dtype is None:
        dtype = np.dtype(type(a)) if isinstance(a, (int, float)) else a.dtype
    return smallest_dtype_for(fill_value, dtype)


_FORCED = {f: np.dtype(bool) for f in ("all", "any", "nanall", "nanany", "allnan", "anynan")}
_FORCED.update({f: np.dtype(np.int64) for f in ("len", "nanlen", "argmax", "argmin", "nanargmax", "nanargmin")})
_FORCED["array"] = np.dtype(object)                                            # utils.py:247-261
_FLOAT_RESULT = {"mean", "var", "std", "nanmean", "nanvar", "nanstd"}           # utils.py:278


def result_dtype(dtype, name, a, n):
    """utils.py:291-344 (``check_dtype``)."""
    if np.isscalar(a) or not a.shape:
        if name not in ("sum", "prod", "len"):
            raise ValueError("scalar inputs are supported only for 'sum', 'prod' and 'len'")
        a_dt = np.dtype(type(a))
    else:
        a_dt = a.dtype
    if dtype is not None:
        if np.issubdtype(dtype, np.bool_) and not ("all" in name or "any" in name):
            raise TypeError(f"function {name} requires a more complex datatype than bool")
        if not np.issubdtype(dtype, np.integer) and name in ("len", "nanlen"):
            raise TypeError(f"function {name} requires an integer datatype")
        return np.dtype(dtype)
    if name in _FORCED:
        return _FORCED[name]
    if name in _FLOAT_RESULT:
        return a_dt if np.issubdtype(a_dt, np.floating) else np.dtype(np.float64)
    if name == "sum":
        if np.issubdtype(a_dt, np.int64):
            return np.dtype(np.int64)
        if np.issubdtype(a_dt, np.integer):
            return smallest_dtype_for(np.iinfo(a_dt).max * n, a_dt)
        if np.issubdtype(a_dt, np.bool_):
            return smallest_dtype_for(n, a_dt)
    return a_dt


def validate_fill(fill_value, dtype, name):
    """utils.py:365-374 (+ check_boolean :182-184)."""
    if name in ("all", "any", "allnan", "anynan"):
        if fill_value not in (0, 1):
            raise ValueError("Value not boolean")
        return
    try:
        dtype.type(fill_value)
    except ValueError:
        raise ValueError(f"fill_value must be convertible into {dtype.type.__name__}")


def _identity_low(fill_value, dtype):      # utils.py:347-353 (``minval``)
    dt = smallest_dtype_for(fill_value, dtype)
    if issubclass(dt.type, np.floating):
        return -np.inf
    if issubclass(dt.type, np.integer):
        return np.iinfo(dt).min
    return np.finfo(dt).min                # bool lands here and raises ValueError, like the reference


def _identity_high(fill_value, dtype):     # utils.py:356-362 (``maxval``)
    dt = smallest_dtype_for(fill_value, dtype)
    if issubclass(dt.type, np.floating):
        return np.inf
    if issubclass(dt.type, np.integer):
        return np.iinfo(dt).max
    return np.finfo(dt).max


# --------------------------------------------------------------------------------------
# 3. Forms 1-5 -> flat labels                                      (utils.py:386-542)
# --------------------------------------------------------------------------------------
def _axis_labels(group_idx, a, axis, size, order):
    """Form 3 label expansion, utils.py:386-432.  Last axis uses the 'offset' recipe
    (C layout regardless of ``order``), any other axis uses ravel_multi_index."""
    nd = a.ndim
    n_groups = int(np.max(group_idx)) + 1 if size is None else size
    out_shape = [n_groups if d == axis else s for d, s in enumerate(a.shape)]
    if axis == nd - 1:                                   # "offset" (utils.py:411-432)
        outer = int(np.prod(a.shape[:-1], dtype=np.int64)) if nd > 1 else 1
        rows = np.arange(outer, dtype=int).reshape(a.shape[:-1] + (1,))
        flat = np.broadcast_to(group_idx, a.shape) + rows * n_groups
    else:                                                # "ravel" (utils.py:394-405)
        grids = []
        for d, s in enumerate(a.shape):
            coord = group_idx if d == axis else np.arange(s)
            shp = [1] * nd
            shp[d] = s
            grids.append(coord.reshape(shp))
        flat = np.ravel_multi_index(grids, out_shape, order=order, mode="raise")
    return flat, out_shape


def flatten_inputs(group_idx, a, size=None, order="C", axis=None, func=None):
    """utils.py:435-542 (``input_validation``): returns
    (flat labels, flat a, flat_size, ndim_idx, out shape, unravel_shape)."""
    if not isinstance(a, (int, float, complex)):
        a = np.asanyarray(a)
    group_idx = np.asanyarray(group_idx)
    if not np.issubdtype(group_idx.dtype, np.integer):
        raise TypeError("group_idx must be of integer type")
    if np.any(group_idx < 0):
        raise ValueError("negative indices not supported")
    nd_idx, nd_a = np.ndim(group_idx), np.ndim(a)

    if axis is None:
        if nd_a > 1:
            raise ValueError("a must be scalar or 1 dimensional, use .ravel to flatten. Alternatively specify axis.")
    elif axis >= nd_a or axis < -nd_a:
        raise ValueError("axis arg too large for np.ndim(a)")
    else:
        if axis < 0:
            axis += nd_a
        if nd_idx > 1:
            raise NotImplementedError("only 1d indexing currently supported with axis arg.")
        if a.shape[axis] != len(group_idx):
            raise ValueError("a.shape[axis] doesn't match length of group_idx.")
        if size is not None and not np.isscalar(size):
            raise NotImplementedError("when using axis arg, size must be None or scalar.")
        form3 = group_idx.ndim == 1 and a.ndim > 1
        in_shape = a.shape if form3 else group_idx.shape
        unravel = in_shape if isinstance(func, str) and "arg" in func else None
        flat, out_shape = _axis_labels(group_idx, a, axis, size, order)
        flat_size = np.prod(out_shape)
        if form3 and not callable(func) and "cum" in func:
            out_shape = in_shape                          # N->N keeps the input shape (utils.py:501-505)
        return flat.ravel(), a.ravel(), flat_size, nd_a, out_shape, unravel

    if nd_idx == 1:
        if size is None:
            size = int(np.max(group_idx)) + 1
        else:
            if not np.isscalar(size):
                raise ValueError("output size must be scalar or None")
            if np.any(group_idx > size - 1):
                raise ValueError(f"one or more indices are too large for size {size}")
        flat_size = size
    else:
        if size is None:
            size = np.max(group_idx, axis=1).astype(int) + 1
        elif np.isscalar(size):
            raise ValueError(f"output size must be of length {len(group_idx)}")
        elif len(size) != len(group_idx):
            raise ValueError(f"{len(size)} sizes given, but {len(group_idx)} output dimensions specified in index")
        group_idx = np.ravel_multi_index(group_idx, size, order=order, mode="raise")
        flat_size = np.prod(size)
    if not (np.ndim(a) == 0 or len(a) == group_idx.size):
        raise ValueError("group_idx and a must be of the same length, or a can be scalar")
    return group_idx, a, flat_size, nd_idx, size, None


# --------------------------------------------------------------------------------------
# 4. Per-function kernels                                   (aggregate_numpy.py:20-274)
# --------------------------------------------------------------------------------------
def _mark_untouched(labels, ret, fill_value):            # aggregate_numpy.py:403-407
    never = np.ones_like(ret, dtype=bool)
    never[labels] = False
    ret[never] = fill_value


def k_sum(labels, a, size, fill_value, dtype=None):      # aggregate_numpy.py:20-39
    dtype = _dtype_for_fill(fill_value, dtype, a)
    if np.ndim(a) == 0:
        ret = np.bincount(labels, minlength=size).astype(dtype, copy=False)
        if a != 1:
            ret *= a
    elif np.iscomplexobj(a):
        raise NotImplementedError("complex input is outside the CUDA path's scope")
    else:
        ret = np.bincount(labels, weights=a, minlength=size).astype(dtype, copy=False)
    if fill_value != 0:
        _mark_untouched(labels, ret, fill_value)
    return ret


def k_prod(labels, a, size, fill_value, dtype=None):     # aggregate_numpy.py:42-48
    dtype = _dtype_for_fill(fill_value, dtype, a)
    ret = np.full(size, fill_value, dtype=dtype)
    if fill_value != 1:
        ret[labels] = 1
    np.multiply.at(ret, labels, a)
    return ret


def k_len(labels, a, size, fill_value, dtype=None):      # aggregate_numpy.py:51-52
    return k_sum(labels, 1, size, fill_value, dtype=int)


def k_last(labels, a, size, fill_value, dtype=None):     # aggregate_numpy.py:55-62
    ret = np.full(size, fill_value, dtype=smallest_dtype_for(fill_value, dtype or a.dtype))
    ret[labels] = a
    return ret


def k_first(labels, a, size, fill_value, dtype=None):    # aggregate_numpy.py:65-69
    ret = np.full(size, fill_value, dtype=smallest_dtype_for(fill_value, dtype or a.dtype))
    ret[labels[::-1]] = a[::-1]
    return ret


def _need_boolean(x):                                    # utils.py:182-184
    if x not in (0, 1):
        raise ValueError("Value not boolean")


def k_all(labels, a, size, fill_value, dtype=None):      # aggregate_numpy.py:72-78
    _need_boolean(fill_value)
    ret = np.full(size, fill_value, dtype=bool)
    if not fill_value:
        ret[labels] = True
    ret[labels.compress(np.logical_not(a))] = False
    return ret


def k_any(labels, a, size, fill_value, dtype=None):      # aggregate_numpy.py:81-87
    _need_boolean(fill_value)
    ret = np.full(size, fill_value, dtype=bool)
    if fill_value:
        ret[labels] = False
    ret[labels.compress(a)] = True
    return ret


def k_min(labels, a, size, fill_value, dtype=None):      # aggregate_numpy.py:90-99
    dtype = smallest_dtype_for(fill_value, dtype or a.dtype)
    top = _identity_high(fill_value, dtype)
    with np.errstate(invalid="ignore"):
        ret = np.full(size, fill_value, dtype=dtype)
    if fill_value != top:
        ret[labels] = top
    with np.errstate(invalid="ignore"):
        np.minimum.at(ret, labels, a)
    return ret


def k_max(labels, a, size, fill_value, dtype=None):      # aggregate_numpy.py:102-111
    dtype = smallest_dtype_for(fill_value, dtype or a.dtype)
    bottom = _identity_low(fill_value, dtype)
    with np.errstate(invalid="ignore"):
        ret = np.full(size, fill_value, dtype=dtype)
    if fill_value != bottom:
        ret[labels] = bottom
    with np.errstate(invalid="ignore"):
        np.maximum.at(ret, labels, a)
    return ret


def _k_arg(extreme, sentinel, labels, a, size, fill_value, dtype, nansqueeze):
    """aggregate_numpy.py:114-139: two passes (group extreme, then first matching index).
    A group holding a NaN has a NaN extreme, matches nothing, and keeps fill_value."""
    probe = np.where(np.isnan(a), sentinel, a) if nansqueeze else a
    best = extreme(labels, probe, size, np.nan)
    hit = a == best[labels]
    ret = np.full(size, fill_value, dtype=dtype)
    (pos,) = hit.nonzero()
    ret[labels[hit][::-1]] = pos[::-1]                    # reversed scatter: first occurrence wins
    return ret


def k_argmax(labels, a, size, fill_value, dtype=int, _nansqueeze=False):
    return _k_arg(k_max, -np.inf, labels, a, size, fill_value, dtype, _nansqueeze)


def k_argmin(labels, a, size, fill_value, dtype=int, _nansqueeze=False):
    return _k_arg(k_min, np.inf, labels, a, size, fill_value, dtype, _nansqueeze)


def k_mean(labels, a, size, fill_value, dtype=np.dtype(np.float64)):   # aggregate_numpy.py:142-161
    if np.ndim(a) == 0:
        raise ValueError("cannot take mean with scalar a")
    if np.iscomplexobj(a):
        raise NotImplementedError("complex input is outside the CUDA path's scope")
    n = np.bincount(labels, minlength=size)
    s = np.bincount(labels, weights=a, minlength=size)
    with np.errstate(divide="ignore", invalid="ignore"):
        ret = s / n
    if not np.isnan(fill_value):
        ret[n == 0] = fill_value
    return ret.astype(dtype, copy=False)


def k_sumofsquares(labels, a, size, fill_value, dtype=np.dtype(np.float64)):   # aggregate_numpy.py:164-172
    ret = np.bincount(labels, weights=a * a, minlength=size)
    if fill_value != 0:
        ret[np.bincount(labels, minlength=size) == 0] = fill_value
    return ret.astype(dtype, copy=False)


def k_var(labels, a, size, fill_value, dtype=np.dtype(np.float64), sqrt=False, ddof=0):   # aggregate_numpy.py:175-195
    if np.ndim(a) == 0:
        raise ValueError("cannot take variance with scalar a")
    n = np.bincount(labels, minlength=size)
    s = np.bincount(labels, weights=a, minlength=size)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = s / n
        n = np.where(n > ddof, n - ddof, 0)
        ret = np.bincount(labels, (a - mean[labels]) ** 2, minlength=size) / n
    if sqrt:
        ret = np.sqrt(ret)
    if not np.isnan(fill_value):
        ret[n == 0] = fill_value
    return ret.astype(dtype, copy=False)


def k_std(labels, a, size, fill_value, dtype=np.dtype(np.float64), ddof=0):    # aggregate_numpy.py:198-199
    return k_var(labels, a, size, fill_value, dtype=dtype, sqrt=True, ddof=ddof)


def k_allnan(labels, a, size, fill_value, dtype=bool):   # aggregate_numpy.py:202-203
    return k_all(labels, np.isnan(a), size, fill_value)


def k_anynan(labels, a, size, fill_value, dtype=bool):   # aggregate_numpy.py:206-207
    return k_any(labels, np.isnan(a), size, fill_value)


def k_sort(labels, a, size=None, fill_value=None, dtype=None, reverse=False):  # aggregate_numpy.py:210-214
    by_group_then_value = np.lexsort((-a if reverse else a, labels))
    slot_of = np.argsort(np.argsort(labels, kind="stable"), kind="stable")
    return a[by_group_then_value][slot_of]


def k_array(labels, a, size, fill_value, dtype=None):    # aggregate_numpy.py:217-227
    if fill_value is not None and not (np.isscalar(fill_value) or len(fill_value) == 0):
        raise ValueError("fill_value must be None, a scalar or an empty sequence")
    by_group = np.argsort(labels, kind="stable")
    n = np.bincount(labels, minlength=size)
    ret = np.asanyarray(np.split(a[by_group], np.cumsum(n)[:-1]), dtype="object")
    if fill_value is None or np.isscalar(fill_value):
        _mark_untouched(labels, ret, fill_value)
    return ret


def k_callable(labels, a, size, fill_value, dtype=None, func=None, **kw):      # aggregate_numpy.py:230-241
    groups = k_array(labels, a, size, ())
    ret = np.full(size, fill_value, dtype=dtype or np.float64)
    for g, members in enumerate(groups):
        if np.ndim(members) == 1 and len(members) > 0:
            ret[g] = func(members)
    return ret


def k_cumsum(labels, a, size, fill_value=None, dtype=None):                    # aggregate_numpy.py:244-266
    by_group = np.argsort(labels, kind="stable")
    back = np.argsort(by_group, kind="stable")
    lab_s, a_s = labels[by_group], a[by_group]
    run = np.cumsum(a_s, dtype=dtype)
    start = k_min(lab_s, np.arange(len(a), dtype=int), size, fill_value=0)[lab_s]
    run -= run[start]                 # large terms first ...
    run += a_s[start]                 # ... then the small one (test_generic.py:577-585)
    return run[back]


def k_nancumsum(labels, a, size, fill_value=None, dtype=None):                 # aggregate_numpy.py:269-274
    return k_cumsum(labels, np.where(np.isnan(a), 0, a), size, fill_value=fill_value, dtype=dtype)


# ---- N->N functions only aggregate_numba has: serial loop semantics -----------------
def _running_loop(labels, a, ret, seen, out, mode, skip_nan):
    """aggregate_numba.py:142-157 with the Prod/Max/Min ``_inner`` bodies (:303-309,
    :369-386) and the N->N ``_outersetter`` (:218-225)."""
    for i in range(len(labels)):
        g = labels[i]
        v = a[i]
        if not (skip_nan and v != v):
            if mode == 0:                     # cumprod
                ret[g] *= v
            elif not seen[g]:                 # first element of the group (also a leading NaN)
                ret[g] = v
                seen[g] = True
            elif mode == 1:                   # cummax
                if ret[g] < v:
                    ret[g] = v
            else:                             # cummin
                if ret[g] > v:
                    ret[g] = v
        out[i] = ret[g]


try:                                          # numba only accelerates the loop; semantics identical
    import numba as _nb
    _running_loop_fast = _nb.njit(cache=False)(_running_loop)
except Exception:                             # pragma: no cover
    _running_loop_fast = None


def _k_running(mode, skip_nan, labels, a, size, fill_value, dtype):
    # ret starts at 1 for cumprod (forced_fill_value, aggregate_numba.py:304) and at
    # fill_value for cummax/cummin (aggregate_numba.py:99-103)
    ret = np.full(size, 1 if mode == 0 else fill_value, dtype=dtype)
    seen = np.zeros(size, dtype=bool)
    out = np.full(len(labels), fill_value, dtype=dtype)
    loop = _running_loop_fast if (_running_loop_fast is not None and len(labels) > 4096) else _running_loop
    loop(np.ascontiguousarray(labels), np.ascontiguousarray(a), ret, seen, out, mode, skip_nan)
    return out


_KERNELS = dict(sum=k_sum, prod=k_prod, len=k_len, last=k_last, first=k_first, all=k_all, any=k_any,
                min=k_min, max=k_max, argmax=k_argmax, argmin=k_argmin, mean=k_mean, var=k_var, std=k_std,
                sumofsquares=k_sumofsquares, allnan=k_allnan, anynan=k_anynan, sort=k_sort, array=k_array,
                cumsum=k_cumsum)
_KERNELS.update({"nan" + k: v for k, v in list(_KERNELS.items()) if k not in _NO_NAN_FORM})
_KERNELS["nancumsum"] = k_nancumsum
_RUNNING = {"cumprod": 0, "cummax": 1, "cummin": 2}


# --------------------------------------------------------------------------------------
# 5. Driver                                                 (aggregate_numpy.py:308-392)
# --------------------------------------------------------------------------------------
def aggregate(group_idx, a, func="sum", size=None, fill_value=0, order="C", dtype=None, axis=None, **kwargs):
    labels, a, flat_size, nd_idx, size, unravel = flatten_inputs(group_idx, a, size=size, order=order,
                                                                 axis=axis, func=func)
    if labels.dtype == np.dtype("uint64"):
        labels = labels.astype(int)                                  # aggregate_numpy.py:324-326
    name = resolve_func(func)
    if not isinstance(name, str):
        ret = k_callable(labels, a, flat_size, fill_value, func=name, dtype=dtype, **kwargs)
    elif (name[3:] if name.startswith("nan") else name) in _RUNNING:
        base = name[3:] if name.startswith("nan") else name
        if np.ndim(a) == 0:
            raise ValueError("scalar inputs are supported only for 'sum', 'prod' and 'len'")
        out_dtype = result_dtype(dtype, name, a, len(labels))         # aggregate_numba.py:68
        validate_fill(fill_value, out_dtype, name)
        ret = _k_running(_RUNNING[base], name.startswith("nan"), labels, a, flat_size, fill_value, out_dtype)
    else:
        if name.startswith("nan"):                                   # aggregate_numpy.py:336-349
            if np.ndim(a) == 0:
                raise ValueError("nan-version not supported for scalar input.")
            if "arg" in name:
                kwargs["_nansqueeze"] = True
            elif "cum" not in name:
                keep = ~np.isnan(a)
                if "len" not in name:
                    a = a[keep]
                labels = labels[keep]
        out_dtype = result_dtype(dtype, name, a, flat_size)
        validate_fill(fill_value, out_dtype, name)
        ret = _KERNELS[name](labels, a, flat_size, fill_value=fill_value, dtype=out_dtype, **kwargs)

    if nd_idx > 1:                                                    # aggregate_numpy.py:358-367
        if unravel is not None:
            blank = ret == fill_value
            ret[blank] = 0
            ret = np.unravel_index(ret, unravel)[axis]
            ret[blank] = fill_value
        ret = ret.reshape(size, order=order)
    return ret


# --------------------------------------------------------------------------------------
# 6. A second CPU baseline: the reference's *numba* strategy (one serial compiled loop per
#    function, aggregate_numba.py:142-157).  Used only by bench.py's cpu_baseline leg so the
#    GPU number is shown next to the faster of the reference's two CPU designs.
# --------------------------------------------------------------------------------------
def _loop_sum(labels, a, ret):
    for i in range(len(labels)):
        ret[labels[i]] += a[i]


def _loop_mean(labels, a, ret, cnt):
    for i in range(len(labels)):
        ret[labels[i]] += a[i]
        cnt[labels[i]] += 1


def _loop_max(labels, a, ret, seen):
    for i in range(len(labels)):
        g = labels[i]
        if not seen[g]:
            ret[g] = a[i]
            seen[g] = True
        elif ret[g] < a[i]:
            ret[g] = a[i]


_serial = {}


def serial_loop_baseline(labels, a, func, size):
    """sum / mean / max the way aggregate_numba runs them (Sum :294-300, Mean :450-462,
    Max :369-376): a single-threaded LLVM-compiled loop.  Requires numba."""
    import numba as nb
    if not _serial:
        _serial.update(sum=nb.njit(_loop_sum), mean=nb.njit(_loop_mean), max=nb.njit(_loop_max))
    if func == "sum":
        ret = np.zeros(size, dtype=a.dtype)
        _serial["sum"](labels, a, ret)
        return ret
    if func == "mean":
        ret = np.zeros(size, dtype=np.float64)
        cnt = np.zeros(size, dtype=np.int64)
        _serial["mean"](labels, a, ret, cnt)
        with np.errstate(divide="ignore", invalid="ignore"):
            out = ret / cnt
        out[cnt == 0] = 0
        return out.astype(a.dtype)
    if func == "max":
        ret = np.zeros(size, dtype=a.dtype)
        seen = np.zeros(size, dtype=np.bool_)
        _serial["max"](labels, a, ret, seen)
        return ret
    raise NotImplementedError(func)
