"""GPU parity of the callers either side of the hot path: unpack / uaggregate, fused multi-statistic
aggregation, device CSR grouping, label utilities, duck arrays.  Checked against the reference's stored answers
(tests/golden/utils.npz) and against the oracle on seeded inputs; integer / index work is bit-exact."""
import numpy as np
import pytest

import utils_golden
from numpy_groupies_b200 import aggregate, aggregate_cuda as ac
from oracle import aggregate_oracle as oracle
from oracle import utils_oracle

pytestmark = pytest.mark.gpu

CASES = utils_golden.cases()


@pytest.mark.parametrize("cid,kind,d", CASES, ids=[f"{c}_{k}" for c, k, _ in CASES])
def test_utils_golden_case(cid, kind, d):
    got = np.asarray(utils_golden.run(kind, d, ac, aggregate))
    ref = d["out"]
    assert got.dtype == ref.dtype, (got.dtype, ref.dtype)
    assert got.shape == ref.shape
    if kind.startswith("uaggregate_") and ref.dtype.kind == "f":
        np.testing.assert_allclose(got, ref, rtol=1e-12)
    else:
        np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("n", [1, 31, 2048, 2049, 1 << 20, (1 << 22) + 77])
def test_step_indices_multi_tile(n):
    rng = np.random.default_rng(n)
    g = np.repeat(rng.integers(0, 1000, n // 3 + 1), rng.integers(1, 6, n // 3 + 1))[:n]
    np.testing.assert_array_equal(ac.step_indices(g), utils_oracle.step_indices(g))
    assert ac.step_count(g) == utils_oracle.step_count(g)


def test_step_indices_empty_and_device_tensor():
    import torch
    assert ac.step_count(np.array([], dtype=np.int64)) == 0
    np.testing.assert_array_equal(ac.step_indices(np.array([], dtype=np.int64)), [0])
    g = torch.tensor([4, 4, 1, 1, 1, 9], device="cuda")
    res = ac.step_indices(g)
    assert res.is_cuda
    np.testing.assert_array_equal(res.cpu().numpy(), [0, 2, 5, 6])


@pytest.mark.parametrize("dtype", [bool, np.int8, np.int32, np.int64, np.uint16])
def test_label_contiguous_large(dtype):
    rng = np.random.default_rng(3)
    n = (1 << 21) + 5
    x = np.repeat(rng.integers(0, 3, n // 4 + 1), rng.integers(1, 8, n // 4 + 1))[:n]
    x = (x != 0) if dtype is bool else x.astype(dtype)
    np.testing.assert_array_equal(ac.label_contiguous_1d(x), utils_oracle.label_contiguous_1d(x))


def test_label_contiguous_rejects_2d():
    with pytest.raises(ValueError):
        ac.label_contiguous_1d(np.zeros((3, 3), dtype=bool))


def test_relabel_and_multi_arange_large():
    rng = np.random.default_rng(5)
    g = rng.integers(0, 300000, 1 << 21) * 2
    got = ac.relabel_groups_unique(g)
    ref = utils_oracle.relabel_groups_unique(g)
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)
    keep = rng.random(int(g.max()) + 1) < 0.3
    np.testing.assert_array_equal(ac.relabel_groups_masked(g, keep), utils_oracle.relabel_groups_masked(g, keep))
    n = rng.integers(0, 9, 300000)
    np.testing.assert_array_equal(ac.multi_arange(n), utils_oracle.multi_arange(n))
    with pytest.raises(ValueError):
        ac.multi_arange(np.zeros((2, 2), dtype=int))


@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int64, np.int16, np.uint8, bool])
def test_unpack_dtypes_and_shapes(dtype):
    rng = np.random.default_rng(11)
    table = (rng.random(5000) * 100).astype(dtype)
    g = rng.integers(0, 5000, (37, 1001))
    got = ac.unpack(g, table)
    assert got.dtype == table.dtype and got.shape == g.shape
    np.testing.assert_array_equal(got, table[g])


def test_unpack_out_of_range_raises_and_device_roundtrip():
    import torch
    with pytest.raises(IndexError):
        ac.unpack(np.array([0, 5]), np.arange(5.0))
    g = torch.randint(0, 100, (1 << 20,), device="cuda")
    t = torch.rand(100, device="cuda", dtype=torch.float64)
    res = ac.unpack(g, t)
    assert res.is_cuda and torch.equal(res, t[g])
    full = ac.uaggregate(g, t[g], func="max", size=100)
    assert full.is_cuda and torch.equal(full, t[g])           # every element equals its group's max here


@pytest.mark.parametrize("adtype", [np.float32, np.float64])
@pytest.mark.parametrize("groups", [100, 5000, 200000])
@pytest.mark.parametrize("fill", [0, -1.5, np.nan])
def test_aggregate_multi_matches_single_calls(adtype, groups, fill):
    rng = np.random.default_rng(groups)
    n = 1 << 20
    g = rng.integers(0, groups, n)
    g[g == 7] = 8                                             # an empty group
    a = rng.standard_normal(n).astype(adtype)
    a[rng.random(n) < 0.05] = np.nan
    funcs = ["sum", "mean", "len", "var", "std", "min", "max", "nansum", "nanmean", "nanvar", "nanlen", "nanmin", "nanmax", "first"]
    # a member the reference would reject (len with fill_value=nan: ValueError) makes the fused call raise like the single call
    singles, rejected = {}, []
    for f in funcs:
        kw = {"ddof": 1} if f.replace("nan", "") in ("var", "std") else {}
        try:
            singles[f] = aggregate(g, a, f, size=groups, fill_value=fill, **kw)
        except ValueError:
            rejected.append(f)
    if rejected:
        with pytest.raises(ValueError):
            ac.aggregate_multi(g, a, funcs, size=groups, fill_value=fill, ddof=1)
        funcs = [f for f in funcs if f not in rejected]
    res = ac.aggregate_multi(g, a, funcs, size=groups, fill_value=fill, ddof=1)
    assert list(res) == funcs
    for f in funcs:
        one = singles[f]
        assert res[f].dtype == one.dtype and res[f].shape == one.shape, f
        if f in ("len", "nanlen", "min", "max", "nanmin", "nanmax", "first"):
            np.testing.assert_array_equal(res[f], one, err_msg=f)
        else:
            # two runs add the same f64 terms in different orders: per-group floor = 1e-12 of the group's own sum of |a|
            x = np.nan_to_num(a.astype(np.float64), nan=0.0)
            cnt = np.maximum(np.bincount(g, minlength=groups), 1).astype(np.float64)
            s1, s2 = np.bincount(g, weights=np.abs(x), minlength=groups), np.bincount(g, weights=x * x, minlength=groups)
            base = f.replace("nan", "")
            floor = {"sum": 1e-12 * s1, "mean": 1e-12 * s1 / cnt, "var": 1e-9 * s2 / cnt, "std": np.sqrt(1e-9 * s2 / cnt)}[base]
            rtol = (2e-5 if base in ("var", "std") else 1e-5) if adtype == np.float32 else 1e-12
            r64, o64 = res[f].astype(np.float64), one.astype(np.float64)
            ok = (np.isnan(r64) & np.isnan(o64)) | (np.abs(r64 - o64) <= rtol * np.abs(o64) + floor + 1e-300)
            assert ok.all(), (f, np.flatnonzero(~ok)[:5], r64[~ok][:5], o64[~ok][:5])


def test_aggregate_multi_against_oracle_and_launch_savings():
    from numpy_groupies_b200 import _abi
    rng = np.random.default_rng(1)
    n, groups = 1 << 21, 1000
    g = rng.integers(0, groups, n)
    a = rng.random(n)
    lib = _abi.load()
    before = lib.agc_launch_count()
    res = ac.aggregate_multi(g, a, ["min", "max", "mean", "sum", "len"], size=groups)
    fused = lib.agc_launch_count() - before
    before = lib.agc_launch_count()
    for f in ("min", "max", "mean", "sum", "len"):
        aggregate(g, a, f, size=groups)
    separate = lib.agc_launch_count() - before
    assert fused < separate
    for f in ("min", "max", "len"):
        np.testing.assert_array_equal(res[f], oracle.aggregate(g, a, func=f, size=groups))
    for f in ("mean", "sum"):
        np.testing.assert_allclose(res[f], oracle.aggregate(g, a, func=f, size=groups), rtol=1e-12)


def test_aggregate_multi_integer_values_and_forms():
    rng = np.random.default_rng(2)
    g = rng.integers(0, 50, (2, 40000))
    a = rng.integers(-1000, 1000, 40000)
    res = ac.aggregate_multi(g, a, ["min", "max", "sum", "mean"], size=(50, 50))
    for f, r in res.items():
        ref = oracle.aggregate(g, a, func=f, size=(50, 50))
        assert r.dtype == ref.dtype and r.shape == ref.shape
        np.testing.assert_allclose(r, ref, rtol=1e-12)
    a3 = rng.random((64, 3000)).astype(np.float32)
    g3 = rng.integers(0, 40, 3000)
    res = ac.aggregate_multi(g3, a3, ["min", "max", "mean", "std"], size=40, axis=1)
    for f, r in res.items():
        ref = oracle.aggregate(g3, a3, func=f, size=40, axis=1)
        assert r.shape == ref.shape
        np.testing.assert_allclose(r, ref, rtol=1e-5)


def test_group_csr_is_the_stable_grouping():
    import torch
    rng = np.random.default_rng(9)
    n, groups = 300000, 777
    g = rng.integers(0, groups, n)
    a = rng.random(n)
    order, offsets, values = ac.group_csr(torch.from_numpy(g).cuda(), torch.from_numpy(a).cuda(), size=groups)
    assert order.is_cuda and offsets.is_cuda and values.is_cuda
    ref_order = np.argsort(g, kind="stable")
    np.testing.assert_array_equal(order.cpu().numpy(), ref_order)
    np.testing.assert_array_equal(offsets.cpu().numpy(), np.concatenate(([0], np.cumsum(np.bincount(g, minlength=groups)))))
    np.testing.assert_array_equal(values.cpu().numpy(), a[ref_order])
    # a device-side per-group function on top of it: segment maxima via the CSR offsets == aggregate(max)
    vals, offs = values.cpu().numpy(), offsets.cpu().numpy()
    seg_max = np.array([vals[offs[i]:offs[i + 1]].max() if offs[i + 1] > offs[i] else 0.0 for i in range(groups)])
    np.testing.assert_array_equal(seg_max, aggregate(g, a, "max", size=groups))


def test_duck_arrays():
    import torch

    class Duck:                                               # minimal array-protocol object (utils.py:684-695)
        def __init__(self, arr):
            self._a = arr
            self.ndim, self.shape, self.dtype = arr.ndim, arr.shape, arr.dtype

        def __array__(self, dtype=None, copy=None):
            return self._a

        def __array_function__(self, *args, **kw):
            return NotImplemented

        def __array_ufunc__(self, *args, **kw):
            return NotImplemented

    g, a = np.array([0, 1, 1, 3]), np.array([1.0, 2.0, 3.0, 4.0])
    assert ac.is_duck_array(Duck(a)) and ac.is_duck_array(a) and ac.is_duck_array(torch.zeros(2)) and not ac.is_duck_array([1, 2])
    np.testing.assert_array_equal(aggregate(Duck(g), Duck(a)), [1.0, 5.0, 0.0, 4.0])


@pytest.mark.parametrize("adtype", [np.float32, np.float64])
@pytest.mark.parametrize("groups", [50, 30000, 400000])
def test_deterministic_mode_same_bits_every_time(adtype, groups):
    """deterministic=True: floating sums / means / variances / products add every group's elements in a fixed order (device
    sort by group, one warp per group).  Same input, same bits -- whatever the kernel tuning -- and still the reference's answer."""
    import torch
    rng = np.random.default_rng(groups)
    n = (1 << 20) + 13
    g = rng.integers(0, groups, n)
    a = (rng.standard_normal(n) * np.exp2(rng.integers(-20, 20, n))).astype(adtype)       # wide range, both signs: order matters
    a[rng.random(n) < 0.02] = np.nan
    tg, ta = torch.from_numpy(g).cuda(), torch.from_numpy(a).cuda()
    for func, kw in (("nansum", {}), ("nanmean", {}), ("nanvar", dict(ddof=1)), ("nanstd", {}), ("sum", {}), ("sumofsquares", {}), ("nanprod", {})):
        vals = ta if func != "nanprod" else (1 + ta * 0).where(torch.isnan(ta), 1 + ta.sign() * 1e-3)
        runs = []
        for fast in (True, False, True):
            runs.append(aggregate(tg, vals, func, size=groups, deterministic=True, fast=fast, **kw))
        assert torch.equal(runs[0].view(torch.int32 if adtype == np.float32 else torch.int64),
                           runs[1].view(torch.int32 if adtype == np.float32 else torch.int64)), func
        assert torch.equal(runs[0].view(torch.int32 if adtype == np.float32 else torch.int64),
                           runs[2].view(torch.int32 if adtype == np.float32 else torch.int64)), func
        if func == "nanprod":
            continue
        with np.errstate(all="ignore"):
            ref = oracle.aggregate(g, a, func=func, size=groups, **kw)
        got = runs[0].cpu().numpy()
        assert got.dtype == ref.dtype
        x = np.nan_to_num(a.astype(np.float64), nan=0.0)
        cnt = np.maximum(np.bincount(g, minlength=groups), 1).astype(np.float64)
        s1, s2 = np.bincount(g, weights=np.abs(x), minlength=groups), np.bincount(g, weights=x * x, minlength=groups)
        base = func.replace("nan", "")
        floor = {"sum": 1e-12 * s1, "mean": 1e-12 * s1 / cnt, "var": 1e-9 * s2 / cnt, "std": np.sqrt(1e-9 * s2 / cnt), "sumofsquares": 1e-12 * s2}[base]
        rtol = 1e-5 if adtype == np.float32 else 1e-12
        ok = (np.isnan(got) & np.isnan(ref)) | (np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= rtol * np.abs(ref) + floor + 1e-300)
        assert ok.all(), (func, np.flatnonzero(~ok)[:5])


def test_deterministic_mode_forms_and_fill():
    rng = np.random.default_rng(4)
    a = rng.standard_normal((50, 2000)).astype(np.float32)
    g = rng.integers(0, 30, 2000)
    for func in ("sum", "mean", "std"):
        got = aggregate(g, a, func, size=31, axis=1, fill_value=-5, deterministic=True)
        ref = oracle.aggregate(g, a, func=func, size=31, axis=1, fill_value=-5)
        assert got.shape == ref.shape and got.dtype == ref.dtype
        np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-6)
    gi = rng.integers(0, 9, (2, 5000))
    ai = rng.integers(-5, 5, 5000)
    np.testing.assert_array_equal(aggregate(gi, ai, "sum", size=(9, 9), deterministic=True), oracle.aggregate(gi, ai, func="sum", size=(9, 9)))


def test_second_device_in_the_same_process():
    """Per-device host state (round-1 advisor finding): the dynamic shared-memory attribute, the SM count and the profile
    events are per device; a process that used cuda:0 must be able to run every large-table kernel on cuda:1."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    rng = np.random.default_rng(3)
    n = 1 << 21
    for groups in (100, 30000, 100000):
        g = rng.integers(0, groups, n)
        a = rng.random(n, dtype=np.float32)
        ref = {f: oracle.aggregate(g, a, func=f, size=groups) for f in ("sum", "mean", "max", "len")}
        for dev in ("cuda:0", "cuda:1", "cuda:0"):
            tg, ta = torch.from_numpy(g).to(dev), torch.from_numpy(a).to(dev)
            for f, r in ref.items():
                got = aggregate(tg, ta, f, size=groups)
                assert str(got.device) == dev
                if f in ("max", "len"):
                    np.testing.assert_array_equal(got.cpu().numpy(), r)
                else:
                    np.testing.assert_allclose(got.cpu().numpy(), r, rtol=1e-5)


@pytest.mark.parametrize("groups", [100, 4000, 24000, 60000])
def test_fused_min_max_shared_memory_kernel(groups):
    """AGC_OP_MINMAX on the shared-memory kernel (two key words per group, regime 30) up to 24.9 K groups, on the general
    kernel beyond: min and max of aggregate_multi are bit-identical to the single calls, NaN / +-inf / empty groups included."""
    import torch
    rng = np.random.default_rng(groups)
    n = (1 << 21) + 5
    g = rng.integers(0, groups, n)
    g[g == 11] = 12                                            # an empty group
    a = rng.standard_normal(n).astype(np.float32)
    a[rng.random(n) < 0.01] = np.nan
    a[3], a[4] = np.inf, -np.inf
    tg, ta = torch.from_numpy(g).cuda(), torch.from_numpy(a).cuda()
    for pre in ("", "nan"):
        res = ac.aggregate_multi(tg, ta, [pre + "min", pre + "max"], size=groups, fill_value=-7)
        if groups <= 24000:
            assert ac.last_regime() in (30, 0), ac.last_regime()
        for f in (pre + "min", pre + "max"):
            one = aggregate(tg, ta, f, size=groups, fill_value=-7)
            assert torch.equal(res[f].view(torch.int32), one.view(torch.int32)), f
            with np.errstate(all="ignore"):
                np.testing.assert_array_equal(res[f].cpu().numpy(), oracle.aggregate(g, a, func=f, size=groups, fill_value=-7))


@pytest.mark.parametrize("func", ["sum", "max", "len", "mean", "cumsum", "first"])
def test_pageable_int64_labels_are_narrowed_on_the_way_to_the_device(func, monkeypatch):
    """Host-array path: pageable int64 / uint64 labels of a Form 1 call with a known size cross PCIe as int32 (narrowed by the
    staging threads, agc_host_narrow_labels).  Same results as the plain upload, and labels the narrowing cannot represent
    are still reported as bad labels -- negative, >= size, >= 2^31."""
    rng = np.random.default_rng(11)
    n, groups = (1 << 21) + 13, 70001                        # 16 MB of int64 labels: above the staging threshold
    gidx = rng.integers(0, groups, n)
    a = rng.random(n).astype(np.float32) if func != "cumsum" else rng.integers(-1000, 1000, n)
    got = aggregate(gidx, a, func=func, size=groups)
    monkeypatch.setattr(ac, "_NARROW_LABELS", False)
    plain = aggregate(gidx, a, func=func, size=groups)
    monkeypatch.setattr(ac, "_NARROW_LABELS", True)
    assert got.dtype == plain.dtype
    if func in ("sum", "mean"):
        np.testing.assert_allclose(got, plain, rtol=1e-6)
        np.testing.assert_allclose(got, oracle.aggregate(gidx, a, func=func, size=groups), rtol=1e-5)
    else:
        np.testing.assert_array_equal(got, plain)
        np.testing.assert_array_equal(got, oracle.aggregate(gidx, a, func=func, size=groups))
    np.testing.assert_array_equal(aggregate(gidx.astype(np.uint64), a, func=func, size=groups), got)
    for bad in (-1, groups, 2**31, 2**31 + 5, -2**40):
        g2 = gidx.copy()
        g2[n // 3] = bad
        with pytest.raises(ValueError):
            aggregate(g2, a, func=func, size=groups)


def test_validated_labels_are_remembered_until_they_change():
    """check=True reads the status record back only until a CUDA label tensor has passed once for this size; any in-place
    write (torch's version counter), another size, or another tensor object -- even at the same address -- is checked again."""
    import torch
    dev = torch.device("cuda")
    n, groups = 1 << 16, 1000
    g = torch.randint(0, groups, (n,), device=dev)
    a = torch.rand(n, device=dev)
    ac._VALID.clear()
    ref = aggregate(g, a, func="sum", size=groups)
    assert id(g) in ac._VALID
    for _ in range(3):                                        # remembered: the same results, no read-back needed
        torch.testing.assert_close(aggregate(g, a, func="sum", size=groups), ref)
    torch.testing.assert_close(aggregate(g, a, func="sum", size=groups, check="always"), ref)
    with pytest.raises(ValueError):                           # the same labels are NOT valid for a smaller size
        aggregate(g, a, func="sum", size=groups // 2)
    g[5] = groups + 7                                         # in-place write: version counter moves, checked again
    with pytest.raises(ValueError):
        aggregate(g, a, func="sum", size=groups)
    with pytest.raises(ValueError):                           # and a tensor that failed is never remembered
        aggregate(g, a, func="max", size=groups)
    g[5] = 3
    torch.testing.assert_close(aggregate(g, a, func="len", size=groups).sum().cpu(), torch.tensor(n))
    del g
    g2 = torch.full((n,), groups, device=dev, dtype=torch.int64)     # a new object (the allocator may hand out the same block)
    with pytest.raises(ValueError):
        aggregate(g2, a, func="sum", size=groups)
    assert aggregate(g2, a, func="sum", size=groups, check=False).shape == (groups,)    # never checked, never remembered
    with pytest.raises(ValueError):
        aggregate(g2, a, func="sum", size=groups)


@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 4099, (1 << 20) + 5])
@pytest.mark.parametrize("shift", [0, 1])
def test_convert_f16_is_numpy_astype(n, shift):
    """agc_convert_f16 (the library's own float16 <-> float32 kernel, no torch arithmetic on the product path): every bit
    pattern class -- normals, subnormals, +-0, +-inf, NaN, round-to-nearest-even ties, overflow to inf -- and pointers that
    are / are not 16-byte aligned (a view that starts one element in)."""
    import torch
    rng = np.random.default_rng(n + shift)
    h = rng.integers(0, 1 << 16, n + shift, dtype=np.uint16).view(np.float16)
    th = torch.from_numpy(h.view(np.int16)).cuda().view(torch.float16)[shift:]
    f = ac._convert_f16(ac._Dev(th, np.float16), False).t
    assert f.dtype == torch.float32
    got32, ref32 = f.cpu().numpy(), h[shift:].astype(np.float32)
    nan32 = np.isnan(ref32)                                 # (NaN payloads are not compared: the hardware conversion quiets a
    np.testing.assert_array_equal(np.isnan(got32), nan32)   #  signalling NaN, numpy's software conversion keeps its bits)
    np.testing.assert_array_equal(got32.view(np.uint32)[~nan32], ref32.view(np.uint32)[~nan32])
    # float32 -> float16: random bit patterns (mostly huge / tiny exponents), values around the float16 grid (ties) and the
    # overflow threshold
    x = rng.integers(0, 1 << 32, n + shift, dtype=np.uint64).astype(np.uint32).view(np.float32)
    if n:
        k = (n + shift) // 2
        grid = rng.integers(0, 0x7c00, k, dtype=np.uint16).view(np.float16).astype(np.float32)
        x[:k] = np.nextafter(grid, np.float32(np.inf)) if shift else grid + np.float32(0.5) * np.spacing(grid.astype(np.float16)).astype(np.float32)
        x[-1] = np.float32(65520.0)
    tx = torch.from_numpy(x.view(np.int32)).cuda().view(torch.float32)[shift:]
    hh = ac._convert_f16(ac._Dev(tx, np.float32), True).t
    assert hh.dtype == torch.float16
    with np.errstate(over="ignore"):
        ref = x[shift:].astype(np.float16)
    got = hh.cpu().numpy()
    nan = np.isnan(ref)
    np.testing.assert_array_equal(np.isnan(got), nan)
    np.testing.assert_array_equal(got.view(np.uint16)[~nan], ref.view(np.uint16)[~nan])


def test_float16_device_tensors_in_and_out():
    """float16 CUDA tensors go through agc_convert_f16 on the way in and on the way out; same answer as the numpy call."""
    import torch
    rng = np.random.default_rng(16)
    g = rng.integers(0, 50, 10_000)
    a = rng.random(10_000).astype(np.float16)
    for func in ("sum", "mean", "max", "cumsum"):
        ref = aggregate(g, a, func=func, size=50)
        got = aggregate(torch.from_numpy(g).cuda(), torch.from_numpy(a).cuda(), func=func, size=50)
        assert got.is_cuda and got.dtype == torch.float16 and ref.dtype == np.float16, (func, got.dtype, ref.dtype)
        np.testing.assert_array_equal(got.cpu().numpy().view(np.uint16), ref.view(np.uint16))
        if func != "cumsum":        # (the reference's float16 cumsum is one global running sum in float16: no yardstick)
            np.testing.assert_allclose(ref.astype(np.float64), oracle.aggregate(g, a, func=func, size=50).astype(np.float64), rtol=2e-3, atol=1e-3)
