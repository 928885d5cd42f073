#!/usr/bin/env python
"""Generate ``tests/golden/`` by running the REAL reference in the build container.

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/*.npz + manifest.json

The reference (``/root/reference/numpy_groupies``) cannot travel to the GPU box, so its
answers are frozen here: for every case below the script calls
``aggregate_numpy.aggregate`` (or ``aggregate_numba.aggregate`` for cumprod / cummax /
cummin, which aggregate_numpy does not implement -- reference tests/__init__.py:33) and
records either the result (dtype, shape, memory order, values) or the exception class.

The cases are the reference's own known-answer vectors (tests/test_generic.py), its
cross-implementation fixture recipe (tests/test_compare.py:57-70), its benchmark recipe
(benchmarks/generic.py:79-88), plus dtype / fill_value / Form 3-4-5 / error grids.

Test infrastructure only -- nothing in the product imports this.
"""
import json
import os
import sys
import warnings

import numpy as np

REF_ROOT = os.environ.get("NPG_REFERENCE", "/root/reference")
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# callables the cases may name (func must survive a JSON round trip)
CALLABLES = {
    "len_gt1": lambda x: len(x) > 1,
    "sum_method": lambda x: x.sum(),
    "ptp": lambda x: x.max() - x.min(),
    "median": np.median,
}


def enc(v):
    """JSON-safe encoding of kwargs values (nan / inf / tuples / dtypes)."""
    if isinstance(v, (np.floating, float)):
        v = float(v)
        if np.isnan(v):
            return {"f": "nan"}
        if np.isinf(v):
            return {"f": "inf" if v > 0 else "-inf"}
        return v
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (tuple, list)):
        return {"t": [enc(x) for x in v]}
    if isinstance(v, np.dtype) or isinstance(v, type):
        return {"dtype": np.dtype(v).name}
    return v


class Book:
    def __init__(self):
        self.arrays = {}
        self.cases = []
        self._seen = {}

    def _store(self, arr):
        arr = np.asarray(arr)
        sig = (arr.dtype.str, arr.shape, arr.tobytes())      # identical inputs are stored once
        key = self._seen.get(sig)
        if key is None:
            key = self._seen[sig] = f"a{len(self.arrays)}"
            self.arrays[key] = arr
        return key

    def add(self, cid, group_idx, a, func="sum", impl="numpy", **kw):
        case = {"id": cid, "impl": impl, "kwargs": {k: enc(v) for k, v in kw.items()}}
        case["gidx"] = self._store(group_idx)
        if np.ndim(a) == 0 and not isinstance(a, np.ndarray):
            case["a_scalar"] = enc(a)
        else:
            case["a"] = self._store(a)
        case["func"] = func if isinstance(func, str) else {"callable": func[1]}
        call_func = func if isinstance(func, str) else CALLABLES[func[1]]
        mod = ref_numba if impl == "numba" else ref_numpy
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                res = mod.aggregate(np.asarray(group_idx).copy(), a.copy() if isinstance(a, np.ndarray) else a,
                                    func=call_func, **kw)
        except Exception as e:                                         # noqa: BLE001
            case["raises"] = type(e).__name__
        else:
            res = np.asarray(res)
            if res.dtype == object:
                lens, vals = [], []
                for item in res:
                    if np.ndim(item) == 0:
                        lens.append(-1)
                    else:
                        lens.append(len(item))
                        vals.append(np.asarray(list(item)))   # rows of a 2-D object array -> numeric
                case["object_lens"] = self._store(np.array(lens, dtype=np.int64))
                case["object_vals"] = self._store(np.concatenate(vals) if vals else np.zeros(0))
            else:
                case["out"] = self._store(res)
                case["dtype"] = res.dtype.name
                case["shape"] = list(res.shape)
                case["fortran"] = bool(np.isfortran(res))
        self.cases.append(case)

    def save(self):
        os.makedirs(OUT_DIR, exist_ok=True)
        np.savez_compressed(os.path.join(OUT_DIR, "cases.npz"), **self.arrays)
        with open(os.path.join(OUT_DIR, "manifest.json"), "w") as f:
            json.dump({"numpy": np.__version__, "n_cases": len(self.cases), "cases": self.cases}, f, indent=0)
        nerr = sum("raises" in c for c in self.cases)
        print(f"wrote {len(self.cases)} cases ({nerr} expected exceptions), {len(self.arrays)} arrays")


FUNCS = ("sum prod min max all any mean std var len argmin argmax anynan allnan cumsum sumofsquares "
         "first last nansum nanprod nanmin nanmax nanmean nanstd nanvar nanlen nanargmin nanargmax "
         "nansumofsquares nanfirst nanlast nanall nanany nancumsum sort").split()
NUMBA_FUNCS = "cumprod cummax cummin nancumprod nancummax nancummin".split()


def build(book):
    rng = np.random.default_rng(100)
    nan = np.nan

    # ---- A. the reference's literal known-answer vectors (tests/test_generic.py) -----------------
    book.add("kat_preserve_missing", np.array([0, 1, 3, 1, 3]), np.arange(101, 106, dtype=int))        # :56-60
    for t in ("int64", "uint32", "uint64", "int32", "int16", "uint8"):                                  # :63-69
        book.add(f"kat_uint_gidx_{t}", np.array([1, 1, 2, 2, 2, 2, 4, 4], dtype=t), np.ones(8), dtype=int)
    g = np.array([2, 2, 4, 4, 4, 7, 7, 7])
    book.add("kat_offset_prod", g, g, func="prod", dtype=int)                                           # :92-95
    for pos in (0, 10, -1):                                                                             # :98-103
        gi = np.arange(5).repeat(5); gi[pos] = -1
        book.add(f"kat_negative_{pos}", gi, np.arange(25))
    book.add("kat_shape_mismatch", np.array((1, 2, 3)), np.array((1, 2)))                               # :111-113
    book.add("kat_lists", np.array([0, 1, 3, 1, 3]), np.arange(101, 106, dtype=int), func="array")      # :116-122
    gi = np.array([0, 1, 2, 3, 3, 3, 3, 4, 5, 5, 5, 6, 5, 4, 3, 8, 8])
    book.add("kat_item_counting", gi, np.arange(gi.size), func=("callable", "len_gt1"))                 # :125-129
    book.add("kat_fill_array_none", np.array([0, 2, 2]), np.arange(3), func="array", fill_value=None)   # :132-143
    book.add("kat_fill_sum_m1", np.array([0, 2, 2]), np.arange(3), func="sum", fill_value=-1)
    gi = np.vstack((np.repeat(np.arange(10), 10), np.repeat([0, 1], 50)))                               # :155-159
    book.add("kat_ndim_gidx_none", gi, 1)
    book.add("kat_ndim_gidx_size", gi, 1, size=(10, 2))
    gi = np.arange(0, 100, 2, dtype=int).repeat(5)
    book.add("kat_len", gi, np.arange(gi.size), func="len")                                             # :178-194
    book.add("kat_len_scalar_sum", gi, 1, func="sum")
    a = rng.random(50); a[::4] = nan; a[::5] = nan
    book.add("kat_nanlen", np.arange(0, 20, 2).repeat(5), a, func="nanlen")                             # :197-208
    for f in ("first", "last"):                                                                         # :211-221
        book.add(f"kat_{f}", gi, np.arange(gi.size), func=f, fill_value=-1)
    for f in ("nanfirst", "nanlast"):                                                                   # :224-241
        for off in (0, 2, 4):
            a = np.arange(gi.size, dtype=float); a[off::5] = nan
            book.add(f"kat_{f}_{off}", gi, a, func=f, fill_value=-1)
    a = rng.random(20)
    for f in ("var", "std"):                                                                            # :244-251
        for ddof in (0, 1, 2):
            book.add(f"kat_ddof_{f}_{ddof}", np.zeros(20, dtype=int), a, func=f, ddof=ddof)
    gi = np.arange(0, 100, dtype=int).repeat(5)
    for f in ("sum", "prod", "mean", "var", "std", "len", "max"):                                       # :254-263
        book.add(f"kat_scalar_{f}", gi, 1, func=f)
    book.add("kat_scalar_sum_f", gi, 2.5, func="sum")
    book.add("kat_scalar_prod_3", gi, 3, func="prod")
    a = rng.random(500); a[::2] = nan
    for f in ("sum", "prod", "mean", "var", "std", "all", "any", "len", "min", "max"):                  # :266-290
        book.add(f"kat_naninput_{f}", gi, a, func=f)
    gi = np.array([0, 0, 0, 0, 3, 3, 3, 3])
    ints = np.array([4, 4, 3, 1, 10, 9, 9, 11])
    with_nan = np.array([4, 4, 3, 1, nan, 1, 2, 3])
    mostly_nan = np.array([4, 4, nan, 1, nan, nan, nan, nan])
    for f in ("argmax", "argmin", "nanargmax", "nanargmin"):                                            # :293-344
        book.add(f"kat_{f}_ints", gi, ints, func=f, fill_value=-1)
        book.add(f"kat_{f}_nan", gi, with_nan, func=f, fill_value=-1)
        book.add(f"kat_{f}_mostlynan", gi, mostly_nan, func=f, fill_value=-1)
    g6 = np.array([0, 1, 2, 0, 1, 2])
    neg = np.array([-np.inf, 0, -np.inf, -np.inf, 0, 0]); pos = -neg                                    # :347-380
    book.add("kat_max_neginf", g6, neg, func="max"); book.add("kat_min_posinf", g6, pos, func="min")
    book.add("kat_argmax_neginf", g6, neg, func="argmax", fill_value=-1)
    book.add("kat_argmin_posinf", g6, pos, func="argmin", fill_value=-1)
    book.add("kat_mean", gi, np.arange(8), func="mean")                                                 # :383-388
    gc = np.array([4, 3, 3, 4, 4, 1, 1, 1, 7, 8, 7, 4, 3, 3, 1, 1])
    ac = np.array([3, 4, 1, 3, 9, 9, 6, 7, 7, 0, 8, 2, 1, 8, 9, 8])
    book.add("kat_cumsum", gc, ac, func="cumsum")                                                       # :391-397
    book.add("kat_nancumsum", [0, 0, 0, 1, 1, 0, 0], [2, 2, nan, 2, 2, 2, 2], func="nancumsum")         # :400-408
    for f in ("cummax", "cummin", "cumprod"):                                                           # :411-417
        book.add(f"kat_{f}", gc, ac, func=f, impl="numba")
    g20 = np.repeat(np.arange(5), 4)
    book.add("kat_list_order", g20, np.arange(20), func="array")                                        # :420-429
    book.add("kat_list_order_rev", g20, np.arange(20)[::-1].copy(), func="array")
    gs = np.array([3, 3, 3, 2, 2, 2, 1, 1, 1]); asr = np.array([3, 2, 1, 3, 4, 5, 5, 10, 1])
    book.add("kat_sort", gs, asr, func="sort"); book.add("kat_sort_rev", gs, asr, func="sort", reverse=True)   # :432-442
    g5 = np.array([1, 2, 2, 0, 1])
    a52 = np.array([[1.0, 2.0], [4.0, 4.0], [5.0, 2.0], [nan, 3.0], [8.0, 7.0]])
    book.add("kat_not_last_axis", g5, a52, func="nanmax", fill_value=nan, axis=0)                       # :501-510
    book.add("kat_callable_axis", np.zeros(10, dtype=int), rng.standard_normal(10), func=("callable", "sum_method"),
             axis=-1, fill_value=0)                                                                     # :513-528
    g12 = np.array([0, 0, 2, 2, 2, 1, 1, 2, 2, 1, 1, 0])
    book.add("kat_argred_nd", g12, np.array([[1] * 12, [1] * 12]), func="argmax", axis=-1)              # :531-538
    book.add("kat_argred_negfill", g12, np.array([[1.0] * 12, [nan] * 12]), func="argmax", axis=-1, fill_value=-1)
    for f in ("nanvar", "nanstd"):                                                                      # :553-574
        for ddof in (0, 1):
            for tag, sel in (("none", None), ("rows", ([1, 4, 5], Ellipsis)), ("cell", (1, (0, 1, 2, 3)))):
                a = np.ones((12, 5))
                if sel is not None:
                    a[sel] = nan
                book.add(f"kat_nanfill_{f}_{ddof}_{tag}", np.zeros(5, dtype=int), a, func=f, axis=-1,
                         fill_value=nan, ddof=ddof)
    book.add("kat_cumsum_accuracy", np.array([0, 0, 0, 0, 1]),
             np.array([0.0, 0.0, 0.0, 3.2768e04, 9.99999975e-06]), func="cumsum", axis=-1)              # :577-585

    # ---- B. test_compare.py fixture recipe (:57-70), scaled to 200 groups (N = 4000) -----------------
    gi = np.repeat(np.arange(200), 2); rs = np.random.RandomState(100); rs.shuffle(gi); gi = np.repeat(gi, 10)
    a = rs.randn(gi.size)
    nana = a.copy(); nana[::3] = nan; nana[: len(gi) // 2] = nan
    somenan = a.copy(); somenan[rs.rand(gi.size) < 0.1] = nan
    for f in FUNCS:
        data = nana if "nan" in f else a
        for fv in (0, 1, nan):
            book.add(f"cmp_{f}_fill{fv}", gi, data, func=f, fill_value=fv)
        book.add(f"cmp_{f}_somenan", gi, somenan, func=f)
    for f in NUMBA_FUNCS:
        data = somenan if "nan" in f else a
        book.add(f"cmp_{f}", gi, data, func=f, impl="numba")
        book.add(f"cmp_{f}_i64", gi, rs.randint(-3, 4, gi.size), func=f, impl="numba")
    for nd in (2, 3):                                                                                   # :147-163
        for order in ("C", "F"):
            gnd = rs.randint(0, 12, size=(nd, 3000)); and_ = rs.rand(3000)
            for f in ("sum", "mean", "max", "len", "argmin", "var", "first"):
                book.add(f"cmp_ndim{nd}_{order}_{f}", gnd, and_, func=f, size=[12] * nd, order=order)
            book.add(f"cmp_ndim{nd}_{order}_sizeNone", gnd, and_, order=order)

    # ---- C. benchmark recipe (benchmarks/generic.py:79-88), N = 5000 -------------------------------------
    rs = np.random.RandomState(100)
    a = rs.random_sample(5000); gi = rs.randint(0, 100, 5000); a[a > 0.8] = 0
    nana = a.copy(); nana[(nana < 0.2) & (nana != 0)] = nan
    for f in FUNCS:
        book.add(f"bench_{f}", gi, nana if "nan" in f else a, func=f, size=110)
    for f in NUMBA_FUNCS:
        book.add(f"bench_{f}", gi, nana if "nan" in f else a, func=f, size=110, impl="numba")

    # ---- D. dtype x func x fill grid (small N, with empty groups) ----------------------------------------
    n = 257
    gi = rs.randint(0, 13, n); gi[gi == 5] = 6                     # group 5 and 13.. stay empty
    dts = ["bool", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64", "float32", "float64", "float16"]
    grid_funcs = ("sum prod min max mean var std len argmin argmax all any first last sumofsquares cumsum "
                  "nansum nanmax nanmean nanlen anynan allnan sort nanargmax nanfirst").split()
    for dt in dts:
        if dt == "bool":
            vals = rs.rand(n) > 0.4
        elif "float" in dt:
            vals = (rs.randn(n) * 10).astype(dt)
        else:
            info = np.iinfo(dt)
            vals = rs.randint(max(info.min, -50), min(info.max, 50) + 1, n).astype(dt)
        for f in grid_funcs:
            book.add(f"dt_{dt}_{f}", gi, vals, func=f, size=16)
            if f in ("sum", "prod", "min", "max", "mean", "first", "last", "len", "argmax", "var", "nansum", "nanmax"):
                book.add(f"dt_{dt}_{f}_fillm1", gi, vals, func=f, size=16, fill_value=-1)
                book.add(f"dt_{dt}_{f}_fillnan", gi, vals, func=f, size=16, fill_value=nan)
                book.add(f"dt_{dt}_{f}_fill1p5", gi, vals, func=f, size=16, fill_value=1.5)
        for udt in ("int32", "float32", "float64", "int64", "bool", "uint8"):
            for f in ("sum", "mean", "max", "len", "any", "prod", "argmax", "cumsum"):
                book.add(f"dt_{dt}_{f}_as_{udt}", gi, vals, func=f, size=16, dtype=udt)
        for f in NUMBA_FUNCS:
            if dt in ("bool", "float16"):
                continue
            book.add(f"dt_{dt}_{f}", gi, vals, func=f, size=16, impl="numba")
    # int sums that need the f64 accumulator to stay exact, and int32 wrap in prod
    big = rs.randint(-2**40, 2**40, n)
    book.add("dt_int64_sum_big", gi, big, func="sum"); book.add("dt_int64_cumsum_big", gi, big, func="cumsum")
    book.add("dt_int32_prod_wrap", gi, rs.randint(-9, 10, n).astype("int32"), func="prod")
    book.add("dt_fill_overflow_int8", gi, vals.astype("int8"), func="max", fill_value=300)
    book.add("dt_fill_overflow_dtype", gi, rs.randn(n), func="max", fill_value=300, dtype="int8")
    book.add("dt_len_float_dtype", gi, rs.randn(n), func="len", dtype="float64")
    book.add("dt_sum_bool_dtype", gi, rs.randn(n), func="sum", dtype="bool")
    book.add("dt_argmax_uint_fillm1", gi, rs.randn(n), func="argmax", fill_value=-1, dtype="uint32")
    book.add("dt_any_fill2", gi, rs.randn(n), func="any", fill_value=2)
    book.add("dt_nanany_fill2", gi, rs.randn(n), func="nanany", fill_value=2)

    # ---- E. Forms 3 / 4 / 5 ------------------------------------------------------------------------------
    form_funcs = ("sum mean max min len var argmax argmin nanargmax cumsum nansum nanmean any first last prod std "
                  "nancumsum sort").split()
    for shape in ((12,), (12, 5), (4, 6, 5), (3, 4, 5, 2)):
        a = rs.randn(*shape)
        a_nan = a.copy(); a_nan.flat[::7] = nan
        for axis in range(-1, len(shape)):
            L = shape[axis]
            gi = rs.randint(0, 4, L)
            for order in ("C", "F"):
                for f in form_funcs:
                    book.add(f"f3_{'x'.join(map(str, shape))}_ax{axis}_{order}_{f}", gi, a_nan if "nan" in f else a,
                             func=f, axis=axis, order=order)
            book.add(f"f3_{'x'.join(map(str, shape))}_ax{axis}_size6", gi, a, func="sum", axis=axis, size=6)
            book.add(f"f3_{'x'.join(map(str, shape))}_ax{axis}_sizesmall", gi, a, func="sum", axis=axis, size=2)
        for f in ("cummax", "cumprod"):
            book.add(f"f3_{'x'.join(map(str, shape))}_{f}", rs.randint(0, 4, shape[-1]), a, func=f, axis=-1, impl="numba")
    a = rs.randn(6, 5)
    book.add("f3_no_axis", rs.randint(0, 3, 6), a)                      # ValueError: a.ndim > 1 without axis
    book.add("f3_axis_too_big", rs.randint(0, 3, 6), a, axis=2)
    book.add("f3_axis_len_mismatch", rs.randint(0, 3, 6), a, axis=1)
    book.add("f3_size_tuple", rs.randint(0, 3, 6), a, axis=0, size=(3, 5))
    book.add("f3_2d_gidx_axis", rs.randint(0, 3, (2, 6)), a, axis=0)
    for nd in (2, 3, 4):
        n = 400
        gnd = np.stack([rs.randint(0, 3 + d, n) for d in range(nd)])
        sz = [3 + d for d in range(nd)]
        v = rs.randn(n); v_nan = v.copy(); v_nan[::5] = nan
        for order in ("C", "F"):
            for f in ("sum", "mean", "max", "len", "argmax", "var", "nansum", "last", "any", "prod", "cumsum", "sort"):
                book.add(f"f4_{nd}d_{order}_{f}", gnd, v_nan if "nan" in f else v, func=f, size=sz, order=order)
            book.add(f"f5_{nd}d_{order}_scalar", gnd, 1, size=sz, order=order)
            book.add(f"f5_{nd}d_{order}_scalar_len", gnd, 3, func="len", size=sz, order=order)
        book.add(f"f4_{nd}d_sizeNone", gnd, v)
        book.add(f"f4_{nd}d_scalar_size", gnd, v, size=5)
        book.add(f"f4_{nd}d_wrong_arity", gnd, v, size=sz + [2])
        book.add(f"f4_{nd}d_out_of_range", gnd, v, size=[2] * nd)
        book.add(f"f4_{nd}d_len_mismatch", gnd, v[:-3], size=sz)

    # ---- F. argument errors / odd inputs -------------------------------------------------------------------
    gi = rs.randint(0, 5, 30); v = rs.randn(30)
    book.add("err_float_gidx", gi.astype(float), v)
    book.add("err_too_large", gi, v, size=3)
    book.add("err_size_tuple_1d", gi, v, size=(5,))
    book.add("err_unknown_func", gi, v, func="bogus")
    book.add("err_nansort", gi, v, func="nansort")
    book.add("err_nanarray", gi, v, func="nanarray")
    book.add("err_rsort", gi, v, func="rsort")
    book.add("err_dsorted", gi, v, func="dsorted")
    book.add("err_nan_scalar", gi, 1, func="nansum")
    book.add("err_fill_nan_int", gi, np.arange(30), func="sum", fill_value=nan, dtype=int)
    book.add("ok_fill_nan_int_a", gi, np.arange(30), func="sum", fill_value=nan)
    book.add("ok_alias_plus", gi, v, func="plus"); book.add("ok_alias_amax", gi, v, func="amax")
    book.add("ok_alias_count", gi, v, func="count"); book.add("ok_alias_times", gi, v, func="times")
    book.add("ok_alias_or", gi, v > 0, func="or"); book.add("ok_alias_and", gi, v > 0, func="and")
    book.add("ok_strided", gi, np.asfortranarray(rs.randn(30, 3))[:, 1], func="sum")
    book.add("ok_list_inputs", [0, 0, 1, 3], [1.5, 2.5, 3.0, 4.0], func="mean")
    book.add("ok_empty_size", np.zeros(0, dtype=int), np.zeros(0), func="sum", size=4)
    book.add("err_empty_nosize", np.zeros(0, dtype=int), np.zeros(0), func="sum")
    book.add("ok_ddof_float", gi, v, func="var", ddof=0.5)
    book.add("ok_ddof_big_fillnan", gi, v, func="var", ddof=50, fill_value=nan)
    book.add("ok_ddof_big_fill0", gi, v, func="std", ddof=50)
    book.add("ok_array_fill_empty", gi, v, func="array", fill_value=(), size=8)
    book.add("ok_array_fill_scalar", gi, v, func="array", fill_value=7, size=8)
    book.add("ok_callable_ptp", gi, v, func=("callable", "ptp"), size=8, fill_value=-1)
    book.add("ok_callable_median_dtype", gi, v, func=("callable", "median"), size=8, dtype="float32")
    book.add("ok_minmax_int_extremes", np.array([0, 0, 1, 1]), np.array([-2**63, 5, 2**63 - 1, 0]), func="max", size=3)
    book.add("ok_minmax_int_extremes_min", np.array([0, 0, 1, 1]), np.array([-2**63, 5, 2**63 - 1, 0]), func="min", size=3)
    book.add("ok_max_negzero", np.array([0, 0, 1, 1]), np.array([0.0, -0.0, -0.0, 0.0]), func="max")
    book.add("ok_sort_nan", gi, np.where(rs.rand(30) < 0.2, nan, v), func="sort")
    book.add("ok_sort_ties", rs.randint(0, 3, 60), rs.randint(0, 4, 60), func="sort")
    book.add("ok_sort_ties_rev", rs.randint(0, 3, 60), rs.randint(0, 4, 60), func="sort", reverse=True)


if __name__ == "__main__":
    if not os.path.isdir(REF_ROOT):
        sys.exit(f"{REF_ROOT} not found: golden vectors can only be (re)generated where the reference is mounted")
    sys.path.insert(0, REF_ROOT)
    from numpy_groupies import aggregate_numba as ref_numba      # noqa: E402
    from numpy_groupies import aggregate_numpy as ref_numpy      # noqa: E402
    book = Book()
    build(book)
    book.save()
