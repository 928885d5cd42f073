"""GPU parity at sizes where the shared-memory fast kernels and the multi-tile scan / sort paths
engage (the golden cases are all small): aggregate_cuda vs the oracle on seeded inputs, plus
size-independent properties at the benchmark's full size."""
import numpy as np
import pytest

from numpy_groupies_b200 import aggregate
from oracle import aggregate_oracle as oracle

pytestmark = pytest.mark.gpu


def _labels(rng, n, groups, dist):
    if dist == "uniform":
        return rng.integers(0, groups, n)
    if dist == "zipf":
        w = 1.0 / np.arange(1, groups + 1) ** 1.1
        ranks = np.searchsorted(np.cumsum(w) / w.sum(), rng.random(n))
        return rng.permutation(groups)[np.minimum(ranks, groups - 1)]
    if dist == "sorted":
        return np.sort(rng.integers(0, groups, n))
    raise ValueError(dist)


FAST_FUNCS = ["sum", "mean", "max", "min", "len", "sumofsquares", "var", "std", "nansum", "nanmean", "nanmax", "nanmin",
              "nanlen", "nanvar"]


def _assert_close_per_group(got, ref, gidx, a, groups, func, rtol):
    """Floating results against the reference: rtol (1e-5 for f32 data, the north_star contract) PLUS a floor that is
    PER GROUP and tied to what any f64 accumulation of that group can lose -- 1e-12 of the group's own sum of |a| (of a^2
    for sums of squares / variances; divided by the count for means) -- never a fraction of the largest group's result."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    g = np.asarray(gidx).ravel()
    x = np.nan_to_num(np.asarray(a, dtype=np.float64).ravel(), nan=0.0, posinf=0.0, neginf=0.0)
    base = func.replace("nan", "")
    cnt = np.maximum(np.bincount(g, minlength=groups).astype(np.float64), 1.0)
    s1 = np.bincount(g, weights=np.abs(x), minlength=groups)
    s2 = np.bincount(g, weights=x * x, minlength=groups)
    if base == "sum":
        floor = 1e-12 * s1
    elif base == "mean":
        floor = 1e-12 * s1 / cnt
    elif base == "sumofsquares":
        floor = 1e-12 * s2
    elif base == "var":
        floor = 1e-9 * s2 / cnt                              # f32 deviations resolve a variance to ~1e-7 of the mean square
    elif base == "std":
        floor = np.sqrt(1e-9 * s2 / cnt)
    else:
        floor = np.zeros(groups)
    floor = floor.reshape(ref.shape) if floor.size == ref.size else floor
    both_nan = np.isnan(got) & np.isnan(ref)
    same_inf = np.isinf(ref) & (got == ref)
    ok = both_nan | same_inf | (np.abs(got - ref) <= rtol * np.abs(ref) + floor + 1e-300)
    if not ok.all():
        i = np.flatnonzero(~ok.ravel())[:5]
        raise AssertionError(f"{func}: {(~ok).sum()} groups off; e.g. groups {i.tolist()}: got {got.ravel()[i]}, ref {ref.ravel()[i]}, "
                             f"allowed {(rtol * np.abs(ref) + floor).ravel()[i]}")



@pytest.mark.parametrize("dist", ["uniform", "zipf", "sorted"])
@pytest.mark.parametrize("groups", [100, 5000, 40000, 300000])
@pytest.mark.parametrize("func", FAST_FUNCS)
def test_f32_reductions_vs_oracle(func, groups, dist):
    rng = np.random.default_rng(hash((func, groups, dist)) % 2**32)
    n = 1 << 20
    gidx = _labels(rng, n, groups, dist)
    a = rng.standard_normal(n).astype(np.float32) * 3 + 1
    if "nan" in func:
        a[rng.random(n) < 0.1] = np.nan
    got = aggregate(gidx, a, func=func, size=groups, fill_value=-1)
    ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-1)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    if func in ("max", "min", "nanmax", "nanmin", "len", "nanlen"):
        np.testing.assert_array_equal(got, ref)            # bit-exact
    else:
        _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=1e-5)     # north_star: rtol 1e-5 for f32


@pytest.mark.parametrize("labels_dtype", [np.int32, np.int64])
def test_fast_path_is_taken_and_matches_general_path(labels_dtype):
    rng = np.random.default_rng(5)
    n, groups = 1 << 21, 1000
    gidx = rng.integers(0, groups, n).astype(labels_dtype)
    a = rng.random(n, dtype=np.float32)
    for func in ("sum", "mean", "max", "min", "len"):
        fast = aggregate(gidx, a, func=func, size=groups)
        slow = aggregate(gidx, a, func=func, size=groups, fast=False)
        if func in ("max", "min", "len"):
            np.testing.assert_array_equal(fast, slow)
        else:
            np.testing.assert_allclose(fast, slow, rtol=1e-6)


@pytest.mark.parametrize("dist", ["uniform", "zipf", "sorted"])
@pytest.mark.parametrize("groups", [20000, 100000])
def test_int32_labels_on_the_large_table_kernels(groups, dist):
    """The int32-label instantiations of the one-word sum / count / filter / hot-key-cache / run kernels."""
    rng = np.random.default_rng(groups + len(dist))
    n = (1 << 21) + 2
    gidx = _labels(rng, n, groups, dist).astype(np.int32)
    a = rng.standard_normal(n).astype(np.float32)
    a[rng.random(n) < 0.03] = np.nan
    for func in ("sum", "nansum", "mean", "nanmean", "max", "nanmin", "len", "nanlen", "var"):
        got = aggregate(gidx, a, func=func, size=groups, fill_value=-1)
        with np.errstate(all="ignore"):
            ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-1)
        assert got.dtype == ref.dtype, func
        if func in ("max", "nanmin", "len", "nanlen"):
            np.testing.assert_array_equal(got, ref, err_msg=func)
        else:
            _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=2e-5 if func == "var" else 1e-5)


def test_fixed_point_sum_is_exact_for_dyadic_data():
    """torch.rand-style values k*2^-24 convert exactly into the fixed-point window, so the fast sum equals
    the exactly-rounded result: compare against an integer sum."""
    rng = np.random.default_rng(11)
    n, groups = 1 << 22, 257
    k = rng.integers(0, 1 << 24, n)
    a = (k * 2.0 ** -24).astype(np.float32)
    gidx = rng.integers(0, groups, n)
    got = aggregate(gidx, a, func="sum", size=groups, dtype=np.float64)
    exact = np.bincount(gidx, weights=k.astype(np.float64), minlength=groups) * 2.0 ** -24
    np.testing.assert_array_equal(got, exact)


def test_wide_dynamic_range_goes_through_the_exact_bypass():
    rng = np.random.default_rng(12)
    n, groups = 1 << 20, 64
    a = (rng.standard_normal(n) * 10.0 ** rng.integers(-30, 30, n)).astype(np.float32)
    a[::1001] = np.inf
    a[5::1003] = -np.inf
    gidx = rng.integers(0, groups, n)
    got = aggregate(gidx, a, func="sum", size=groups, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        ref = oracle.aggregate(gidx, a, func="sum", size=groups, dtype=np.float64)
    np.testing.assert_allclose(got, ref, rtol=1e-9, equal_nan=True)


def test_bad_labels_raise_on_the_fast_path():
    rng = np.random.default_rng(13)
    n, groups = 1 << 20, 100
    gidx = rng.integers(0, groups, n)
    a = rng.random(n, dtype=np.float32)
    for bad in (-1, groups):
        g2 = gidx.copy(); g2[n // 3] = bad
        with pytest.raises(ValueError):
            aggregate(g2, a, func="sum", size=groups)


@pytest.mark.parametrize("func", ["cumsum", "cummax", "cummin", "cumprod", "sort"])
@pytest.mark.parametrize("dist", ["uniform", "sorted"])
def test_int64_n_to_n_bit_exact(func, dist):
    rng = np.random.default_rng(21)
    n, groups = 1 << 20, 3000
    gidx = _labels(rng, n, groups, dist)
    a = rng.integers(-2**31, 2**31, n) if func != "cumprod" else rng.integers(-3, 4, n)
    got = aggregate(gidx, a, func=func, size=groups)
    ref = oracle.aggregate(gidx, a, func=func, size=groups)
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("groups", [3, 700, 40000, 2000000])
@pytest.mark.parametrize("n", [(1 << 22) + 777, 1 << 21, 2048 * 3 + 5])
def test_sorted_label_scans_chained_kernel(n, groups):
    """Sorted int64 labels go through the single-pass chained scan: one descriptor per 256-element warp tile, look-back over
    32 tiles at a time.  3 groups = chains thousands of tiles long (every look-back walks several windows), 2 * 10^6 groups
    = a head in almost every tile; ragged N exercises the tail kernel behind the full tiles.  Bit-exact for the integers."""
    rng = np.random.default_rng(n % 1000 + groups)
    gidx = np.sort(rng.integers(0, groups, n))
    a = rng.integers(-2**40, 2**40, n)
    for func in ("cumsum", "cummax", "cummin"):
        np.testing.assert_array_equal(aggregate(gidx, a, func=func, size=groups), oracle.aggregate(gidx, a, func=func, size=groups))
    small = rng.integers(-2, 3, n)
    np.testing.assert_array_equal(aggregate(gidx, small, func="cumprod", size=groups), oracle.aggregate(gidx, small, func="cumprod", size=groups))
    g32 = gidx.astype(np.int32)                              # int32 labels: 32-byte rows, the 32-byte swizzle
    for func in ("cumsum", "cummin"):
        np.testing.assert_array_equal(aggregate(g32, a, func=func, size=groups), oracle.aggregate(gidx, a, func=func, size=groups))
    u = rng.integers(0, 2**63, n).astype(np.uint64) * np.uint64(2) + np.uint64(1)          # values above 2^63: the order-preserving key matters
    for func in ("cummax", "cummin"):
        got = aggregate(gidx, u, func=func, size=groups)
        ref = oracle.aggregate(gidx, u, func=func, size=groups)
        assert got.dtype == ref.dtype
        np.testing.assert_array_equal(got, ref)
    # f64 running sums: the reference takes ONE np.cumsum over the whole array and subtracts each group's offset
    # (aggregate_numpy.py:258-265), so ITS rounding error follows the global running sum (~n / 2 here), not the group's;
    # the scan adds per group in a tree, the reference sequentially over up to 1.4 * 10^6 terms.  Allowed: eps * (global sum)
    # absolute plus rtol 1e-10 (sequential f64 summation of 10^6 terms of one sign is itself only good to ~1e-11).
    atol = 4 * 2.2e-16 * n
    f = rng.random(n)
    np.testing.assert_allclose(aggregate(gidx, f, func="cumsum", size=groups), oracle.aggregate(gidx, f, func="cumsum", size=groups), rtol=1e-10, atol=atol)
    f[rng.random(n) < 0.1] = np.nan
    np.testing.assert_allclose(aggregate(gidx, f, func="nancumsum", size=groups), oracle.aggregate(gidx, f, func="nancumsum", size=groups), rtol=1e-10, atol=atol)
    # one inversion in the middle of the array: the chained kernel must notice and the sort-based path must take over
    g2 = gidx.copy()
    g2[n // 2], g2[n // 2 + 1] = groups - 1, 0
    np.testing.assert_array_equal(aggregate(g2, a, func="cumsum", size=groups), oracle.aggregate(g2, a, func="cumsum", size=groups))


@pytest.mark.parametrize("dtype,lo,hi", [(np.int64, 7, 8), (np.int64, -100, 100), (np.int64, -35000, 35000), (np.int64, -2**31, 2**31),
                                         (np.int64, -2**62, 2**62), (np.int32, -2**20, 2**20), (np.uint16, 0, 65536), (np.int8, -128, 128), (np.uint64, 2**63 - 5, 2**63 + 500)])
@pytest.mark.parametrize("reverse", [False, True])
def test_sort_with_range_compressed_value_keys(dtype, lo, hi, reverse):
    """func='sort' on integer values sorts (key - min) and skips, on the device, the radix passes above bits(max - min):
    0 passes (all equal), 1, 3 (odd: the positions change buffers once more), 4, 8; ties keep their order (stable).
    reverse=True sorts -a as numpy computes it: the negation wraps in the array's own width, so 0 of an unsigned type and the
    most negative value of a signed one stay in front (uint16 with zeros, int8 with -128)."""
    rng = np.random.default_rng(5)
    n, groups = (1 << 20) + 321, 3000
    gidx = rng.integers(0, groups, n)
    if dtype == np.uint64:
        a = (rng.integers(0, hi - lo, n).astype(np.uint64) + np.uint64(lo))
    else:
        a = rng.integers(lo, hi, n).astype(dtype)
    got = aggregate(gidx, a, func="sort", size=groups, reverse=reverse)
    ref = oracle.aggregate(gidx, a, func="sort", size=groups, reverse=reverse)
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("func", ["first", "last", "argmax", "argmin", "nanargmax", "any", "all", "prod", "nanfirst"])
def test_other_reductions_f64(func):
    rng = np.random.default_rng(31)
    n, groups = 1 << 20, 20000
    gidx = rng.integers(0, groups, n)
    a = rng.standard_normal(n)
    if func == "prod":
        a = 1 + a * 1e-3
    if "nan" in func:
        a[rng.random(n) < 0.1] = np.nan
    if func in ("any", "all"):
        a = (a > -1.5)
    got = aggregate(gidx, a, func=func, size=groups)
    ref = oracle.aggregate(gidx, a, func=func, size=groups)
    assert got.dtype == ref.dtype
    if func == "prod":
        np.testing.assert_allclose(got, ref, rtol=1e-12)
    else:
        np.testing.assert_array_equal(got, ref)


def test_config3_var_std_nanmean_f64():
    """BASELINE config 3 at reduced N: f64, 10 % NaN, many groups, ddof=1, rtol 1e-12."""
    rng = np.random.default_rng(41)
    n, groups = 1 << 21, 100000
    gidx = rng.integers(0, groups, n)
    a = rng.random(n)
    a[rng.random(n) < 0.1] = np.nan
    for func, kw in (("nanvar", dict(ddof=1)), ("nanstd", dict(ddof=1)), ("nanmean", {}), ("var", dict(ddof=1))):
        got = aggregate(gidx, a, func=func, size=groups, fill_value=np.nan, **kw)
        with np.errstate(all="ignore"):
            ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=np.nan, **kw)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-15, equal_nan=True)


def test_form3_and_form4_large():
    rng = np.random.default_rng(51)
    a = rng.random((64, 4096), dtype=np.float32)
    gidx = rng.integers(0, 37, 4096)
    for func in ("sum", "mean", "max"):
        got = aggregate(gidx, a, func=func, axis=1, size=37)
        ref = oracle.aggregate(gidx, a, func=func, axis=1, size=37)
        np.testing.assert_allclose(got, ref, rtol=1e-5)
    g2 = rng.integers(0, 256, (2, 1 << 20))
    v = rng.random(1 << 20, dtype=np.float32)
    for func in ("sum", "mean"):
        got = aggregate(g2, v, func=func, size=(256, 256))
        ref = oracle.aggregate(g2, v, func=func, size=(256, 256))
        np.testing.assert_allclose(got, ref, rtol=1e-5)


def test_full_size_properties_device_resident():
    """N = 2^28 on the device: linearity / conservation properties that need no CPU oracle."""
    import torch
    n, groups = 1 << 28, 1000
    g = torch.Generator(device="cuda").manual_seed(100)
    gidx = torch.randint(0, groups, (n,), generator=g, device="cuda")
    a = torch.rand(n, generator=g, device="cuda")
    s = aggregate(gidx, a, func="sum", size=groups)
    c = aggregate(gidx, a, func="len", size=groups)
    m = aggregate(gidx, a, func="mean", size=groups)
    mx = aggregate(gidx, a, func="max", size=groups)
    assert int(c.sum()) == n                                            # counts conserve N (bit-exact)
    torch.testing.assert_close(s.double().sum(), a.double().sum(), rtol=1e-6, atol=0)   # sum of sums
    torch.testing.assert_close(m.double(), (s.double() / c.double()), rtol=1e-6, atol=0)
    assert float(mx.max()) == float(a.max())                            # max of maxes
    s2 = aggregate(gidx, 2 * a, func="sum", size=groups)                # linearity (exact: power-of-two scaling)
    assert torch.equal(s2, 2 * s)


# ---------------------------------------------------------------------------------------------------------
# tables larger than one SM's shared memory: every kernel of agc_dispatch.cuh forced in turn
# ---------------------------------------------------------------------------------------------------------
CLUSTER_FUNCS = ["sum", "mean", "max", "min", "len", "sumofsquares", "var", "nansum", "nanmean", "nanmax", "nanlen"]


@pytest.fixture(params=[0, 3, 4], ids=["cache", "partial", "runs"])
def force_large(request):
    """Force the kernel used for tables larger than one SM's shared memory: 0 = k_cache, 3 = partial k_fast,
    4 = k_runs (segmented warp reduce; on random labels it degenerates into one RED per element)."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    ac.tuning(large=request.param)
    yield ac, request.param
    ac.tuning(large=1)


@pytest.mark.parametrize("dist", ["uniform", "zipf", "sorted"])
@pytest.mark.parametrize("groups", [40000, 100000, 230000])
@pytest.mark.parametrize("func", CLUSTER_FUNCS)
def test_large_table_kernels_vs_oracle(func, groups, dist, force_large):
    """Every large-table kernel forced onto every label distribution (the probe would route them apart): same
    tolerances as the other f32 reductions; the regime word proves which kernel ran."""
    ac, strategy = force_large
    rng = np.random.default_rng(hash((func, groups, dist, "cl")) % 2**32)
    n = (1 << 21) + 3                                       # several phases per CTA + a ragged tail
    gidx = _labels(rng, n, groups, dist)
    a = rng.standard_normal(n).astype(np.float32) * 3 + 1
    if "nan" in func:
        a[rng.random(n) < 0.1] = np.nan
    a[::4099] *= 1e12                                       # out-of-window values take the exact bypass
    got = aggregate(gidx, a, func=func, size=groups, fill_value=-1)
    regime = ac.last_regime()
    ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-1)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    words = {"sum": 3, "nansum": 3, "sumofsquares": 3, "mean": 3, "nanmean": 3, "var": 3}.get(func, 1)   # fill=-1: sum+count
    if groups * words * 4 > 226 * 1024 and not (func == "var" and groups * 4 <= 195 * 1024):
        # table larger than one SM's shared memory (var up to 49.9 K groups: the regime word is the second pass's, k_fast)
        assert regime // 10 == {0: 2, 3: 1, 4: 5}[strategy], regime
    if func in ("max", "min", "nanmax", "len", "nanlen"):
        np.testing.assert_array_equal(got, ref)
    else:
        _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=2e-5 if func == "var" else 1e-5)


def test_probe_routes_uniform_and_skewed_labels():
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(77)
    n, groups = 1 << 21, 100000
    a = rng.random(n, dtype=np.float32)
    seen = {}
    for dist in ("uniform", "zipf", "sorted"):
        gidx = _labels(rng, n, groups, dist)
        got = aggregate(gidx, a, func="sum", size=groups)
        seen[dist] = ac.last_regime()
        np.testing.assert_allclose(got, oracle.aggregate(gidx, a, func="sum", size=groups), rtol=1e-5)
    assert seen["uniform"] == 15 and seen["zipf"] == 20 and seen["sorted"] == 50, seen


def test_large_table_kernels_full_size_agree():
    """N = 2^28, 10^5 groups, device resident: the partial-table k_fast and k_runs against k_cache (bit-exact for
    max / len; sums are exact fixed point / f64 partials up to the final f64 RED order)."""
    import torch
    from numpy_groupies_b200 import aggregate_cuda as ac
    n, groups = 1 << 28, 100000
    g = torch.Generator(device="cuda").manual_seed(100)
    gidx = torch.randint(0, groups, (n,), generator=g, device="cuda")
    a = torch.rand(n, generator=g, device="cuda")
    res = {}
    try:
        for mode in (3, 4, 0):
            ac.tuning(large=mode)
            for func in ("sum", "mean", "max", "min", "len"):
                res[mode, func] = aggregate(gidx, a, func=func, size=groups, check="always")   # (the regime travels in the status record)
                assert (ac.last_regime() // 10) == {3: 1, 4: 5, 0: 2}[mode], (mode, func, ac.last_regime())
    finally:
        ac.tuning(large=1)
    for mode in (3, 4):
        for func in ("max", "min", "len"):
            assert torch.equal(res[mode, func], res[0, func]), (mode, func)
        for func in ("sum", "mean"):
            torch.testing.assert_close(res[mode, func], res[0, func], rtol=1e-6, atol=0)
    assert int(res[3, "len"].sum()) == n


@pytest.mark.parametrize("func", ["sum", "mean", "max", "min", "nansum", "nanmean", "nanmax", "var", "sumofsquares"])
@pytest.mark.parametrize("shape,groups", [((37, 8192), 100), ((5, 4, 4096), 1000), ((2, 65536), 9000)])
def test_form3_last_axis_fast_path(func, shape, groups):
    """Form 3 on the last axis of a C-contiguous f32 array: the row-block kernel (regime 40+mode) vs the oracle."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(hash((func, shape, groups)) % 2**32)
    a = (rng.standard_normal(shape) * 2 + 0.5).astype(np.float32)
    if "nan" in func:
        a[rng.random(shape) < 0.1] = np.nan
    a.reshape(-1)[::5003] *= 1e10
    gidx = rng.integers(0, groups, shape[-1])
    got = aggregate(gidx, a, func=func, axis=-1, size=groups, fill_value=-3)
    assert ac.last_regime() // 10 == 4, ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, axis=-1, size=groups, fill_value=-3)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    if func in ("max", "min", "nanmax"):
        np.testing.assert_array_equal(got, ref)
    else:
        outer = int(np.prod(shape[:-1]))
        flat = (np.arange(outer)[:, None] * groups + gidx[None, :]).ravel()       # group of every element, C order as the result
        _assert_close_per_group(got, ref, flat, a, outer * groups, func, rtol=1e-5)


def test_form3_fast_path_bad_label_and_int32_labels():
    rng = np.random.default_rng(61)
    a = rng.random((16, 8192), dtype=np.float32)
    gidx = rng.integers(0, 50, 8192).astype(np.int32)
    np.testing.assert_allclose(aggregate(gidx, a, func="sum", axis=1, size=50), oracle.aggregate(gidx, a, func="sum", axis=1, size=50), rtol=1e-5)
    g2 = gidx.copy(); g2[4000] = 50
    with pytest.raises(ValueError):
        aggregate(g2, a, func="sum", axis=1, size=50)


# ---------------------------------------------------------------------------------------------------------
# one-word sums with carry forwarding (FM_SUM1, regime 15): plain sums on tables larger than 28 K groups
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dist", ["uniform", "zipf", "sorted"])
@pytest.mark.parametrize("groups", [40000, 100000, 230000])
@pytest.mark.parametrize("func", ["sum", "nansum", "sumofsquares"])
def test_one_word_sums_vs_oracle(func, groups, dist):
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(hash((func, groups, dist, "s1")) % 2**32)
    n = (1 << 21) + 1
    gidx = _labels(rng, n, groups, dist)
    a = rng.standard_normal(n).astype(np.float32) * 3 + 1
    if "nan" in func:
        a[rng.random(n) < 0.1] = np.nan
    a[::4099] *= 1e12
    try:
        ac.tuning(large=3)
        got = aggregate(gidx, a, func=func, size=groups)
        assert ac.last_regime() == 15, ac.last_regime()
    finally:
        ac.tuning(large=1)
    ref = oracle.aggregate(gidx, a, func=func, size=groups)
    assert got.dtype == ref.dtype
    _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=1e-5)


@pytest.mark.parametrize("signed", [False, True])
@pytest.mark.parametrize("groups", [50000, 100000])
def test_one_word_sums_are_exact_for_dyadic_data(groups, signed):
    """Values k * 2^-24 (optionally centred on 0): every carry / borrow out of the 32-bit word must be forwarded
    exactly, so the f64 result equals the integer sum -- few groups so that each word wraps thousands of times."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(91 + signed)
    n = 1 << 22
    k = rng.integers(0, 1 << 24, n) - ((1 << 23) if signed else 0)
    a = (k * 2.0 ** -24).astype(np.float32)
    gidx = rng.integers(0, 300, n)                         # 300 busy groups inside a large table
    gidx[::1000] = rng.integers(0, groups, gidx[::1000].size)
    try:
        ac.tuning(large=3)
        got = aggregate(gidx, a, func="sum", size=groups, dtype=np.float64)
        assert ac.last_regime() == 15
    finally:
        ac.tuning(large=1)
    exact = np.bincount(gidx, weights=k.astype(np.float64), minlength=groups) * 2.0 ** -24
    np.testing.assert_array_equal(got, exact)


@pytest.mark.parametrize("func", ["cumsum", "cummax", "cummin"])
def test_unsorted_n_to_n_multi_million(func):
    """n > 2^22 unsorted labels: many tiles per radix pass and per segmented scan (bit-exact against the oracle)."""
    rng = np.random.default_rng(71)
    n, groups = (1 << 22) + 5, 50000
    gidx = rng.integers(0, groups, n)
    a = rng.integers(-2**31, 2**31, n)
    got = aggregate(gidx, a, func=func, size=groups)
    ref = oracle.aggregate(gidx, a, func=func, size=groups)
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)
    af = rng.standard_normal(n)
    if func == "cumsum":
        np.testing.assert_allclose(aggregate(gidx, af, func=func, size=groups), oracle.aggregate(gidx, af, func=func, size=groups), rtol=1e-9, atol=1e-9)


def test_sorted_labels_full_size_run_kernel():
    """N = 2^28 sorted labels, 10^5 groups, device resident: the probe routes to k_runs (regime 50+mode); results agree
    with the hot-key-cache kernel (bit-exact for max / len, 1e-6 for the f64-accumulated sums)."""
    import torch
    from numpy_groupies_b200 import aggregate_cuda as ac
    n, groups = 1 << 28, 100000
    g = torch.Generator(device="cuda").manual_seed(100)
    gidx = torch.sort(torch.randint(0, groups, (n,), generator=g, device="cuda")).values
    a = torch.rand(n, generator=g, device="cuda")
    res = {}
    try:
        for mode in (1, 0):
            ac.tuning(large=mode)
            for func in ("sum", "mean", "max", "len"):
                res[mode, func] = aggregate(gidx, a, func=func, size=groups, check="always")
                assert (ac.last_regime() // 10) == (5 if mode == 1 else 2), (mode, func, ac.last_regime())
    finally:
        ac.tuning(large=1)
    for func in ("max", "len"):
        assert torch.equal(res[1, func], res[0, func]), func
    for func in ("sum", "mean"):
        torch.testing.assert_close(res[1, func], res[0, func], rtol=1e-6, atol=0)


# ---------------------------------------------------------------------------------------------------------
# small tables on the general path (agc_small.cuh, regime 60): any dtype / Form, additive + multiplicative ops
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("groups", [7, 300, 4000, 12000])
@pytest.mark.parametrize("func,dtype", [("sum", np.float64), ("nansum", np.float64), ("mean", np.float64), ("nanmean", np.float64),
                                        ("var", np.float64), ("nanstd", np.float64), ("sumofsquares", np.float64), ("prod", np.float64),
                                        ("sum", np.int64), ("sum", np.int16), ("sum", np.uint8), ("mean", np.int32), ("var", np.int64),
                                        ("prod", np.int64), ("sumofsquares", np.int32), ("var", np.float32), ("prod", np.float32),
                                        ("sum", np.bool_)])
def test_small_table_general_kernel(func, dtype, groups):
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(hash((func, np.dtype(dtype).name, groups)) % 2**32)
    n = (1 << 18) + 11
    gidx = rng.integers(0, groups, n)
    if np.dtype(dtype).kind == "f":
        a = rng.standard_normal(n).astype(dtype)
        if func == "prod":
            a = (1 + a * 1e-3).astype(dtype)
        if "nan" in func:
            a[rng.random(n) < 0.1] = np.nan
    elif np.dtype(dtype).kind == "b":
        a = rng.random(n) < 0.3
    elif func == "prod":
        a = rng.integers(-2, 3, n).astype(dtype)
    else:
        info = np.iinfo(dtype)
        lim = 30 if func == "sumofsquares" else 1000         # keep int32 sums of squares inside the dtype (overflow is garbage in the reference)
        a = rng.integers(max(info.min, -lim), min(info.max, lim) + 1, n).astype(dtype)
    got = aggregate(gidx, a, func=func, size=groups, fill_value=-1 if np.dtype(dtype).kind != "b" else 0, ddof=1) if "var" in func or "std" in func \
        else aggregate(gidx, a, func=func, size=groups)
    if np.dtype(dtype).kind != "f" or dtype != np.float32 or func == "prod":
        assert ac.last_regime() == 60, ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-1, ddof=1) if "var" in func or "std" in func \
            else oracle.aggregate(gidx, a, func=func, size=groups)
    assert got.dtype == ref.dtype and got.shape == ref.shape
    if got.dtype.kind == "f":
        rtol = 1e-5 if dtype == np.float32 else 1e-12
        if dtype == np.float32 and func.replace("nan", "") in ("var", "std"):
            rtol = 2e-5
        if dtype == np.float32 and func.replace("nan", "") == "prod":
            rtol = 2e-4        # the REFERENCE multiplies ~37 000 factors per group in float32 (np.multiply.at): ITS rounding, ~sqrt(n) ulp
        _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=rtol)
    else:
        np.testing.assert_array_equal(got, ref)


def test_small_table_general_kernel_forms_3_4():
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(83)
    a = rng.standard_normal((300, 1000))                      # f64: general path
    gidx = rng.integers(0, 9, 1000)
    for func in ("sum", "mean", "var"):
        got = aggregate(gidx, a, func=func, axis=1, size=9)
        assert ac.last_regime() == 60
        np.testing.assert_allclose(got, oracle.aggregate(gidx, a, func=func, axis=1, size=9), rtol=1e-12, atol=1e-12)
    g2 = rng.integers(0, 40, (2, 1 << 18))
    v = rng.integers(-50, 50, 1 << 18)
    got = aggregate(g2, v, func="sum", size=(40, 40))
    assert ac.last_regime() == 60
    np.testing.assert_array_equal(got, oracle.aggregate(g2, v, func="sum", size=(40, 40)))


def test_len16_counter_overflow_is_forwarded_exactly():
    """k_len16: two 16-bit counters per word; hot groups wrap the 15-bit range hundreds of times.  Neighbouring groups
    (same word) must not disturb each other and nothing may be lost."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(97)
    n, groups = (1 << 25) + 3, 100000          # ~100 K adds per hot group and CTA: each wraps the 15-bit range three times
    gidx = rng.integers(0, groups, n)
    hot = rng.random(n)
    gidx[hot < 0.45] = 4242            # even group: low half of its word
    gidx[(hot >= 0.45) & (hot < 0.9)] = 4243   # odd neighbour: high half of the same word
    a = rng.random(n, dtype=np.float32)
    a[rng.random(n) < 0.05] = np.nan
    try:
        ac.tuning(large=3)
        got = aggregate(gidx, a, func="len", size=groups)
        assert ac.last_regime() == 16, ac.last_regime()
        np.testing.assert_array_equal(got, np.bincount(gidx, minlength=groups))
        got = aggregate(gidx, a, func="nanlen", size=groups)
        np.testing.assert_array_equal(got, np.bincount(gidx[~np.isnan(a)], minlength=groups))
        m = aggregate(gidx, a, func="nanmean", size=groups)
        with np.errstate(all="ignore"):
            ref = oracle.aggregate(gidx, a, func="nanmean", size=groups)
        np.testing.assert_allclose(m, ref, rtol=1e-5, equal_nan=True)
        # worst case for the read-check-repair protocol: every thread of every CTA hammers the same half
        one = np.full(1 << 24, 77777, dtype=np.int64)
        one[::4097] = 77776
        got = aggregate(one, 1.0, func="len", size=groups)
        assert ac.last_regime() == 16, ac.last_regime()
        np.testing.assert_array_equal(got, np.bincount(one, minlength=groups))
        # groups beyond the counter table (size > ~99 800) go straight to the global counts
        big = 130000
        gb = rng.integers(0, big, 1 << 22)
        got = aggregate(gb, 1.0, func="len", size=big)
        assert ac.last_regime() == 16, ac.last_regime()
        np.testing.assert_array_equal(got, np.bincount(gb, minlength=big))
    finally:
        ac.tuning(large=1)


@pytest.mark.parametrize("func", ["sum", "mean", "nansum", "var"])
@pytest.mark.parametrize("binades", [4, 30])
def test_hot_key_cache_side_accumulator(func, binades):
    """Skewed labels on a large table (k_cache, regime 20 / 21): when the label probe also finds > 0.1 % of the sampled
    values below the fixed-point window (status[7]), every cache slot carries an f64 side accumulator; narrow data runs
    the layout without it.  Both must match the oracle."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(binades + len(func))
    n, groups = (1 << 21) + 7, 100000
    gidx = _labels(rng, n, groups, "zipf")
    a = (np.exp2(rng.uniform(-binades / 2, binades / 2, n)) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    if "nan" in func:
        a[rng.random(n) < 0.05] = np.nan
    got = aggregate(gidx, a, func=func, size=groups)
    if func != "var":
        assert ac.last_regime() // 10 == 2, ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups)
    _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=2e-5 if func == "var" else 1e-5)


@pytest.mark.parametrize("groups", [100, 5000, 12000])
@pytest.mark.parametrize("func", ["sum", "mean", "nansum", "nanmean", "sumofsquares", "var"])
def test_small_table_side_accumulator_for_wide_dynamic_range(func, groups):
    """Data over 40 binades (and both signs): most values miss the CTA's fixed-point window and go through the f64 side
    accumulator in shared memory instead of global REDs; inf / NaN travel the same way."""
    rng = np.random.default_rng(groups + len(func))
    n = (1 << 20) + 5
    gidx = rng.integers(0, groups, n)
    a = (np.exp2(rng.uniform(-20, 20, n)) * rng.choice([-1.0, 1.0], n)).astype(np.float32)
    if "nan" in func:
        a[rng.random(n) < 0.05] = np.nan
    if func in ("sum", "nansum"):
        a[gidx == 7] = 1.5
        a[np.flatnonzero(gidx == 7)[:1]] = np.inf            # one group with an infinity
    got = aggregate(gidx, a, func=func, size=groups)
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups)
    assert got.dtype == ref.dtype
    _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=2e-5 if func == "var" else 1e-5)


@pytest.mark.parametrize("dist", ["uniform", "zipf"])
@pytest.mark.parametrize("func", ["mean", "nanmean", "var", "nanstd", "sum", "nansum"])
def test_one_word_sum_and_count_table(func, dist):
    """16.6 K < size <= 24.9 K groups: mean / var (and sums with a fill value, which need the touched flags) keep a
    one-word carry-forwarding sum and a u32 count per group in one pass (k_fast<FM_SUMCNT1>, regime 19)."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(hash((func, dist, "c1")) % 2**32)
    n, groups = (1 << 21) + 3, 20000
    gidx = _labels(rng, n, groups - 50, dist)                  # the last 50 groups stay empty
    a = (rng.standard_normal(n) * 3 + 1).astype(np.float32)
    a[::7919] *= 1e8                                          # far outside the exact window: f64 bypass
    a[::4001] = 0.0
    if "nan" in func:
        a[rng.random(n) < 0.1] = np.nan
    got = aggregate(gidx, a, func=func, size=groups, fill_value=-2)
    if func in ("mean", "nanmean", "sum", "nansum"):
        assert ac.last_regime() == 19, ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-2)
    assert got.dtype == ref.dtype
    np.testing.assert_allclose(got, ref, rtol=2e-5 if "var" in func or "std" in func else 1e-5, atol=1e-30, equal_nan=True)


@pytest.mark.parametrize("groups", [99776, 99778, 150000, 400000])
@pytest.mark.parametrize("func", ["max", "nanmin"])
def test_max_min_filter_kernel_partial_table(func, groups):
    """Sizes around and beyond what the 16-bit filter table covers (99 776 groups in 195 KB): the groups past the table
    always pass the filter and go to the global table directly."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(groups + len(func))
    n = (1 << 21) + 1
    gidx = rng.integers(0, groups, n)
    a = rng.standard_normal(n).astype(np.float32)
    a[rng.random(n) < 0.02] = np.nan
    got = aggregate(gidx, a, func=func, size=groups, fill_value=7)
    assert ac.last_regime() == (17 if func == "max" else 18), ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=7)
    np.testing.assert_array_equal(got, ref)
    # a hot group past the table: the probe sees the skew; the whole table is covered up to 113 K groups (tie-testing
    # filter kernel), beyond that the hot-key cache takes over -- a hot group never sits among uncovered ones
    gidx[rng.random(n) < 0.03] = groups - 3
    got = aggregate(gidx, a, func=func, size=groups, fill_value=7)
    want = (17 if func == "max" else 18) if groups <= 113000 else (22 if func == "max" else 23)
    assert ac.last_regime() == want, ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=7)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("dist", ["uniform", "zipf", "sorted"])
@pytest.mark.parametrize("func", ["max", "min", "nanmax", "nanmin"])
def test_max_min_filter_kernel(func, dist):
    """56 K < size <= 113 K groups: k_maxfilter (regime 17 / 18), bit-exact incl. NaN propagation, +-inf, -0.0 vs 0.0
    ordering aside, all-NaN groups and never-touched groups."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(hash((func, dist, "flt")) % 2**32)
    n, groups = (1 << 21) + 2, 100000
    gidx = _labels(rng, n, groups - 100, dist)                 # the last 100 groups stay empty -> fill_value
    a = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 4, n)).astype(np.float32)
    a[rng.random(n) < 0.05] = np.nan
    a[::5003] = np.inf
    a[7::5003] = -np.inf
    a[gidx == gidx[0]] = np.nan                               # one group with nothing but NaN
    got = aggregate(gidx, a, func=func, size=groups, fill_value=-5)
    if dist == "sorted":
        assert ac.last_regime() // 10 == 5, ac.last_regime()      # labels in runs: the probe hands them to k_runs
    else:
        assert ac.last_regime() == (17 if "max" in func else 18), ac.last_regime()
    with np.errstate(all="ignore"):
        ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-5)
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)


# ---------------------------------------------------------------------------------------------------------
# the headline size itself: N = 2^30 needs MORE THAN ONE launch of the shared-memory kernels (a CTA takes at most 2^22
# elements per launch: fixed-point / u32-count headroom), so the second chunk's path (labels / values offset by the first
# chunk, tables not re-initialised) is only reached here.  Checked against an independent device reference (torch f64
# index_add_ / bincount / scatter_reduce -- in TESTS only), bit-exact for len / max, 1e-6 for sums and means.
# ---------------------------------------------------------------------------------------------------------
def test_second_launch_chunk_at_n_2p30():
    import torch
    from numpy_groupies_b200 import _abi
    if torch.cuda.get_device_properties(0).total_memory < 60e9:
        pytest.skip("needs ~40 GB of device memory")
    lib = _abi.load()
    dev = torch.device("cuda")
    n = 1 << 30
    gen = torch.Generator(device=dev).manual_seed(5)
    a = torch.rand(n, generator=gen, device=dev)
    for func, groups in (("sum", 20000), ("mean", 10000), ("len", 20000), ("max", 20000), ("nansum", 3000)):
        lab = torch.randint(0, groups, (n,), generator=gen, device=dev)
        before = lib.agc_launch_count()
        got = aggregate(lab, a, func=func, size=groups)
        launches = lib.agc_launch_count() - before
        if func in ("sum", "mean"):                         # one CTA per SM (table > 97 KB): 148 x 2^22 elements per launch < 2^30
            assert launches >= 4, (func, launches)          # init + two (or more) accumulate launches + finalise
        sums = torch.zeros(groups, dtype=torch.float64, device=dev)
        cnt = torch.zeros(groups, dtype=torch.int64, device=dev)
        mx = torch.full((groups,), float("-inf"), device=dev)
        for s in range(0, n, 1 << 27):
            e = s + (1 << 27)
            sums.index_add_(0, lab[s:e], a[s:e].double())
            cnt += torch.bincount(lab[s:e], minlength=groups)
            mx.scatter_reduce_(0, lab[s:e], a[s:e], "amax")
        if func == "len":
            assert torch.equal(got, cnt)
        elif func == "max":
            assert torch.equal(got, mx)
        elif func == "mean":
            torch.testing.assert_close(got.double(), sums / cnt.double(), rtol=1e-6, atol=0)
        else:
            torch.testing.assert_close(got.double(), sums, rtol=1e-6, atol=0)
        del lab


def test_nanvar_second_pass_keeps_nan_deviations():
    """A +-inf in a group makes its mean inf / NaN: the squared deviation of the group's ORDINARY elements is then NaN and
    must propagate (only original NaNs are skipped, aggregate_numpy.py:336-349 then :184-186) -- on every kernel the dispatcher
    may pick for the second pass (round-1 advisor finding: the shared-memory second pass dropped them)."""
    rng = np.random.default_rng(77)
    n = 1 << 20
    for groups in (500, 20000, 40000):                      # k_accumulate_small / two-word k_fast / one-word k_fast second passes
        gidx = rng.integers(0, groups, n)
        a = rng.standard_normal(n).astype(np.float32)
        a[rng.random(n) < 0.05] = np.nan
        hot = np.flatnonzero(gidx == 3)
        a[hot[0]] = np.inf
        a[hot[1]] = 1.0
        for func in ("nanvar", "nanstd"):
            got = aggregate(gidx, a, func=func, size=groups)
            with np.errstate(all="ignore"):
                ref = oracle.aggregate(gidx, a, func=func, size=groups)
            assert np.isnan(ref[3]) and np.isnan(got[3]), (groups, func, got[3], ref[3])
            ok = ~np.isnan(ref)
            np.testing.assert_allclose(got[ok], ref[ok], rtol=2e-5, atol=1e-30)


# ---------------------------------------------------------------------------------------------------------
# k_team (agc_team.cuh, opt-in with tuning(team=2|4)): a thread-block cluster owns the table, foreign elements travel
# as st.async records over DSMEM.  Same exactness as the one-word sums; counts ride as 16-bit read-check-repair counters.
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("labels_dtype", [np.int64, np.int32])
@pytest.mark.parametrize("groups", [60000, 100001, 230000])
@pytest.mark.parametrize("team", [2, 4])
def test_team_kernel_vs_oracle(team, groups, labels_dtype):
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(team * 1000 + groups)
    n = (1 << 21) + 3
    gidx = rng.integers(0, groups, n).astype(labels_dtype)
    a = rng.standard_normal(n).astype(np.float32) * 3 + 1
    a[::4099] *= 1e12                                       # out-of-window values: exact f64 bypass
    a[rng.random(n) < 0.05] = np.nan
    a[5] = np.inf
    try:
        ac.tuning(team=team, team_maxu=90)
        for func in ("nansum", "nanmean", "sum", "mean", "nanvar"):
            got = aggregate(gidx, a, func=func, size=groups, fill_value=-1)
            if func != "nanvar":
                assert ac.last_regime() in (70, 71), (func, ac.last_regime())
            with np.errstate(all="ignore"):
                ref = oracle.aggregate(gidx, a, func=func, size=groups, fill_value=-1)
            assert got.dtype == ref.dtype and got.shape == ref.shape
            _assert_close_per_group(got, ref, gidx, a, groups, func, rtol=2e-5 if "var" in func else 1e-5)
        bad = gidx.copy(); bad[n // 2] = groups
        with pytest.raises(ValueError):
            aggregate(bad, a, func="sum", size=groups)
    finally:
        ac.tuning(team=0, team_maxu=60)


def test_team_kernel_is_exact_for_dyadic_data_and_counts_hot_groups():
    """Integer-valued f32 data: every partial sum is exact, so the result must EQUAL the integer sum whatever route an
    element took (own table, record to a peer, overflow of a record buffer, carry out of a 32-bit word).  One hot group takes
    a third of the elements: its 16-bit counter wraps through the repair path thousands of times."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(9)
    n, groups = (1 << 22) + 1, 70000
    gidx = rng.integers(0, groups, n)
    gidx[rng.random(n) < 0.33] = 12345
    a = rng.integers(-2000, 2000, n).astype(np.float32)
    ref_sum = np.bincount(gidx, weights=a.astype(np.float64), minlength=groups)
    ref_cnt = np.bincount(gidx, minlength=groups)
    try:
        for team in (2, 4):
            ac.tuning(team=team, team_maxu=90, large=3)     # large=3: no probe, the uniform-label kernels run whatever the skew
            got = aggregate(gidx, a, func="sum", size=groups)
            assert ac.last_regime() == 70
            np.testing.assert_array_equal(got.astype(np.float64), ref_sum.astype(np.float32).astype(np.float64))
            mean = aggregate(gidx, a, func="mean", size=groups)
            assert ac.last_regime() == 71
            with np.errstate(all="ignore"):
                np.testing.assert_array_equal(mean, (ref_sum / ref_cnt).astype(np.float32))
    finally:
        ac.tuning(team=0, team_maxu=60, large=1)


@pytest.mark.parametrize("func", ["sum", "mean", "max", "len"])
def test_partial_table_kernel_on_fewer_ctas(func):
    """agc_tuning key 9: the partial-table kernel on any number of CTAs (grid-stride loops, one table flush per CTA) gives the
    same result -- bit for bit on dyadic data, whose sums are exact whatever the order."""
    from numpy_groupies_b200 import aggregate_cuda as ac
    rng = np.random.default_rng(99)
    n, groups = (1 << 22) + 13, 100_000
    a = (rng.integers(0, 1 << 24, n) * 2.0 ** -24).astype(np.float32)
    gidx = rng.integers(0, groups, n)
    try:
        ac.tuning(large=3)
        ref = aggregate(gidx, a, func=func, size=groups)
        regime = ac.last_regime()
        for grid in (1, 37, 100, 1000):                    # (more CTAs than SMs are ignored: one per SM)
            ac.tuning(part_grid=grid)
            got = aggregate(gidx, a, func=func, size=groups)
            assert ac.last_regime() == regime
            np.testing.assert_array_equal(got, ref)
    finally:
        ac.tuning(large=1, part_grid=0)
    want = oracle.aggregate(gidx, a, func=func, size=groups)
    if func in ("max", "len"):
        np.testing.assert_array_equal(ref, want)
    else:
        np.testing.assert_allclose(ref, want, rtol=1e-6)
