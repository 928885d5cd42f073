"""Host-side semantics of ``aggregate``: name resolution, Forms 1-5, result dtype and
fill_value rules -- restated from the reference's shared rules module so that
``aggregate_cuda`` answers exactly like ``aggregate_numpy`` does (same dtypes, shapes and
exception classes).  Pure numpy on *metadata only*: nothing here touches the N elements.

Reference locations (``numpy_groupies/utils.py`` unless noted):
  aliasing / get_func ........ :58-179          input_validation .......... :435-542
  minimum_dtype .............. :205-244          check_dtype ............... :291-344
  check_fill_value ........... :365-374          per-kernel dtype widening . aggregate_numpy.py:20-139
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Tuple

import numpy as np

# ------------------------------------------------------------------------------------------------
# function names
# ------------------------------------------------------------------------------------------------
NO_NAN_FORM = frozenset(("sort", "rsort", "array", "allnan", "anynan"))                 # utils.py:59

_SYNONYMS = {
    "any": ("or", any, np.any), "all": ("and", all, np.all),
    "sum": ("add", "plus", sum, np.add, np.sum), "len": ("count", len),
    "prod": ("multiply", "product", "times", np.multiply, np.prod),
    "max": ("amax", "maximum", max, np.amax, np.max, np.maximum),
    "min": ("amin", "minimum", min, np.amin, np.min, np.minimum),
    "array": ("split", "splice", slice, list, np.array, np.asarray),
    "sort": ("sorted", "asort", "asorted", "rsorted", "dsort", sorted, np.sort),
    "rsort": ("dsorted",),
    "argmax": (np.argmax,), "argmin": (np.argmin,), "mean": (np.mean,), "std": (np.std,), "var": (np.var,),
    "cumsum": (np.cumsum,), "cumprod": (np.cumprod,),
    "nansum": (np.nansum,), "nanprod": (np.nanprod,), "nanmean": (np.nanmean,), "nanvar": (np.nanvar,),
    "nanstd": (np.nanstd,), "nanmax": (np.nanmax,), "nanmin": (np.nanmin,), "nanargmax": (np.nanargmax,),
    "nanargmin": (np.nanargmin,), "nancumsum": (np.nancumsum,),
    "first": (), "last": (), "allnan": (), "anynan": (), "sumofsquares": (), "cummax": (), "cummin": (),
}


def _alias_table():
    tab = {}
    for name, others in _SYNONYMS.items():
        tab[name] = name
        tab.update((o, name) for o in others)
    for name in list(_SYNONYMS):
        if name not in NO_NAN_FORM and not name.startswith("nan"):
            tab["nan" + name] = "nan" + name
    return tab


ALIASES = _alias_table()

# what aggregate_cuda implements: everything aggregate_numpy has (aggregate_numpy.py:277-305)
# plus cumprod/cummax/cummin from aggregate_numba (aggregate_numba.py:491-500); `rsort` exists only
# as an alias target in the reference and no implementation provides it.
_BASE = ("sum prod len last first all any min max argmax argmin mean std var sumofsquares allnan anynan "
         "sort array cumsum cumprod cummax cummin").split()
SUPPORTED = frozenset(_BASE) | frozenset("nan" + f for f in _BASE if f not in NO_NAN_FORM)


def canonical_func(func):
    """str / builtin / numpy callable -> canonical name; other callables are returned unchanged."""
    try:
        name = ALIASES[func]
    except KeyError:
        if callable(func):
            return func
        raise ValueError(f"func {func!r} is neither a valid function string nor a callable object")
    if name not in SUPPORTED:
        raise NotImplementedError("No such function available")
    return name


# ------------------------------------------------------------------------------------------------
# dtypes and fill values
# ------------------------------------------------------------------------------------------------
_WIDER_INT = {"bool": "int8", "uint8": "int16", "int8": "int16", "uint16": "int32", "int16": "int32",
              "uint32": "int64", "int32": "int64"}
_WIDER_FLOAT = {"float16": "float32", "float32": "float64", "float64": "complex64", "complex64": "complex128"}


def _holds(x, dt):
    try:
        with np.errstate(invalid="ignore"):
            y = np.array(x).astype(dt)
    except (ValueError, OverflowError, RuntimeWarning):
        return False
    return y == x or np.isnan(x)


def minimum_dtype(x, dtype=np.bool_):
    """Narrowest dtype at least as wide as ``dtype`` that represents ``x`` (NaN always fits)."""
    dt = np.dtype(dtype)
    inexact = np.issubdtype(dt, np.inexact)
    chain = _WIDER_FLOAT if inexact else _WIDER_INT
    while not _holds(x, dt):
        nxt = chain.get(dt.name)
        if nxt is None:
            if inexact:
                raise ValueError(f"Can not determine dtype of {x!r}")
            return np.dtype(np.float32)
        dt = np.dtype(nxt)
    return dt


_BOOL_RESULT = frozenset(("all", "any", "nanall", "nanany", "allnan", "anynan"))
_INT_RESULT = frozenset(("len", "nanlen", "argmax", "argmin", "nanargmax", "nanargmin"))
_FLOAT_RESULT = frozenset(("mean", "var", "std", "nanmean", "nanvar", "nanstd"))


def scalar_dtype(a):
    return np.dtype(type(a))


def result_dtype(user_dtype, name, a_dtype, a_is_scalar, n):
    """``check_dtype``: the dtype handed to the per-function kernel."""
    if a_is_scalar and name not in ("sum", "prod", "len"):
        raise ValueError("scalar inputs are supported only for 'sum', 'prod' and 'len'")
    if user_dtype is not None:
        if np.issubdtype(user_dtype, np.bool_) and not ("all" in name or "any" in name):
            raise TypeError(f"function {name} requires a more complex datatype than bool")
        if not np.issubdtype(user_dtype, np.integer) and name in ("len", "nanlen"):
            raise TypeError(f"function {name} requires an integer datatype")
        return np.dtype(user_dtype)
    if name in _BOOL_RESULT:
        return np.dtype(bool)
    if name in _INT_RESULT:
        return np.dtype(np.int64)
    if name == "array":
        return np.dtype(object)
    if name in _FLOAT_RESULT:
        return a_dtype if np.issubdtype(a_dtype, np.floating) else np.dtype(np.float64)
    if name == "sum":
        if np.issubdtype(a_dtype, np.int64):
            return np.dtype(np.int64)
        if np.issubdtype(a_dtype, np.integer):
            return minimum_dtype(np.iinfo(a_dtype).max * n, a_dtype)
        if np.issubdtype(a_dtype, np.bool_):
            return minimum_dtype(n, a_dtype)
    return a_dtype


def validate_fill(fill_value, dtype, name):
    """``check_fill_value`` (+ the ``check_boolean`` calls inside _all/_any)."""
    base = name[3:] if name.startswith("nan") else name
    if base in ("all", "any", "allnan", "anynan"):        # nanall / nanany reach check_boolean inside _all/_any
        if fill_value not in (0, 1):
            raise ValueError("Value not boolean")
        if name == base:
            return
    try:
        dtype.type(fill_value)
    except ValueError:
        raise ValueError(f"fill_value must be convertible into {dtype.type.__name__}")


def kernel_out_dtype(name, dtype, a_dtype, fill_value):
    """The dtype the reference's kernel finally allocates (its own widening for fill_value)."""
    base = name[3:] if name.startswith("nan") else name
    if base in ("sum", "prod"):
        return minimum_dtype(fill_value, dtype)                      # minimum_dtype_scalar, dtype is never None here
    if base == "len":
        return minimum_dtype(fill_value, np.dtype(int))              # _len hard-codes dtype=int
    if base in ("min", "max", "first", "last"):
        return minimum_dtype(fill_value, dtype or a_dtype)
    if base in ("all", "any", "allnan", "anynan"):
        return np.dtype(bool)
    if base == "sort":
        return a_dtype
    return np.dtype(dtype)


def probe_fill(fill_value, dtype):
    """Raise exactly what ``np.full(size, fill_value, dtype)`` would (OverflowError / ValueError)."""
    np.full((), fill_value, dtype=dtype)


def extreme_identity_check(name, fill_value, out_dtype):
    """_min/_max call maxval/minval, which raise ValueError for bool (np.finfo(bool), utils.py:347-362)."""
    dt = minimum_dtype(fill_value, out_dtype)
    if not (issubclass(dt.type, np.floating) or issubclass(dt.type, np.integer)):
        np.finfo(dt)


# ------------------------------------------------------------------------------------------------
# Forms 1-5 -> how the device derives the flat group of element i
# ------------------------------------------------------------------------------------------------
@dataclass
class IndexPlan:
    form: str                         # "flat" | "axis" | "multi"
    flat_size: int                    # number of flat groups
    out_shape: Optional[Tuple[int, ...]]   # shape handed to reshape(..., order) ; None for 1-D results
    ndim_idx: int
    n: int                            # number of elements
    a_shape: Tuple[int, ...] = ()
    axis: int = 0
    dims: Tuple[int, ...] = ()
    strides: Tuple[int, ...] = ()
    axis_size: int = 0
    arg_stride: int = 0               # arg* on Form 3: along-axis index = (i // arg_stride) % arg_len
    arg_len: int = 0
    needs_size: bool = False          # size=None: max label must be measured on the device first
    extra: dict = field(default_factory=dict)


def _strides(shape, order):
    s, acc = [0] * len(shape), 1
    rng = range(len(shape) - 1, -1, -1) if order == "C" else range(len(shape))
    for d in rng:
        s[d] = acc
        acc *= int(shape[d])
    return tuple(s)


def plan_index(gidx_shape, gidx_is_int, a_shape, a_is_scalar, size, order, axis, func):
    """Shape-level part of ``input_validation``.  ``size`` may contain None entries to be resolved
    from the labels (see ``resolve_size``)."""
    if not gidx_is_int:
        raise TypeError("group_idx must be of integer type")
    nd_idx, nd_a = len(gidx_shape), len(a_shape)
    if order not in ("C", "F"):
        raise ValueError("order must be 'C' or 'F'")
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
        if a_shape[axis] != gidx_shape[0]:
            raise ValueError("a.shape[axis] doesn't match length of group_idx.")
        if size is not None and not np.isscalar(size):
            raise NotImplementedError("when using axis arg, size must be None or scalar.")
        form3 = nd_a > 1
        n = int(np.prod(a_shape, dtype=np.int64))
        plan = IndexPlan(form="axis", flat_size=0, out_shape=None, ndim_idx=nd_a, n=n, a_shape=tuple(a_shape),
                         axis=axis, needs_size=size is None)
        plan.extra = dict(order=order, user_size=size, form3=form3,
                          keep_shape=form3 and not callable(func) and "cum" in func,
                          arg=isinstance(func, str) and "arg" in func)
        if size is not None:
            finish_axis_plan(plan, int(size))
        return plan

    if nd_idx == 1:
        if size is not None and not np.isscalar(size):
            raise ValueError("output size must be scalar or None")
        n = int(gidx_shape[0])
        plan = IndexPlan(form="flat", flat_size=0 if size is None else size, out_shape=None, ndim_idx=1, n=n,
                         needs_size=size is None)
    else:
        if size is None:
            pass
        elif np.isscalar(size):
            raise ValueError(f"output size must be of length {gidx_shape[0]}")
        elif len(size) != gidx_shape[0]:
            raise ValueError(f"{len(size)} sizes given, but {gidx_shape[0]} output dimensions specified in index")
        n = int(gidx_shape[1])
        plan = IndexPlan(form="multi", flat_size=0, out_shape=None, ndim_idx=nd_idx, n=n, needs_size=size is None)
        plan.extra = dict(order=order, rows=int(gidx_shape[0]))
        if size is not None:
            finish_multi_plan(plan, [int(s) for s in size])
    if not (a_is_scalar or (nd_a >= 1 and a_shape[0] == n)):
        raise ValueError("group_idx and a must be of the same length, or a can be scalar")
    return plan


def finish_axis_plan(plan, n_groups):
    """Form 3 strides.  The last axis uses the reference's 'offset' recipe, which lays groups out in C
    order whatever ``order`` says (utils.py:411-432, 495); other axes use ravel_multi_index(order)."""
    a_shape, axis, order = plan.a_shape, plan.axis, plan.extra["order"]
    out_shape = tuple(n_groups if d == axis else int(s) for d, s in enumerate(a_shape))
    layout = "C" if axis == len(a_shape) - 1 else order
    plan.dims = tuple(int(s) for s in a_shape)
    plan.strides = _strides(out_shape, layout)
    plan.axis_size = n_groups
    plan.flat_size = int(np.prod(out_shape, dtype=np.int64))
    plan.out_shape = tuple(a_shape) if plan.extra["keep_shape"] else out_shape
    if plan.extra["arg"]:
        plan.arg_len = int(a_shape[axis])
        plan.arg_stride = int(np.prod(a_shape[axis + 1:], dtype=np.int64))
    return plan


def finish_multi_plan(plan, sizes):
    plan.dims = tuple(sizes)
    plan.strides = _strides(sizes, plan.extra["order"])
    plan.flat_size = int(np.prod(sizes, dtype=np.int64))
    plan.out_shape = tuple(sizes)
    return plan
