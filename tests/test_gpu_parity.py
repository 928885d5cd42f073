"""GPU parity: aggregate_cuda (through the C ABI) against the reference's recorded answers
(tests/golden) and against the oracle on seeded inputs.  Run with -m gpu on the B200 box."""
import warnings

import numpy as np
import pytest

from golden_io import Golden, check_result
from numpy_groupies_b200 import aggregate

pytestmark = pytest.mark.gpu
G = Golden()

# Deliberate, documented divergences from the reference (DESIGN.md "Parity notes"):
#  * Form 3 on the last axis with labels >= size: the reference's 'offset' recipe never checks and
#    silently bleeds into the next row (utils.py:411-432); aggregate_cuda raises ValueError.
#  * float cumsum when `a` holds NaN/inf: the reference's global np.cumsum leaks them into later
#    groups; aggregate_cuda scans per group like aggregate_numba.
#  * prod with floating `a` and an integer dtype: ufunc.at truncates after every multiply.
def _divergent(case):
    cid = case["id"]
    if "_sizesmall" in cid and "raises" not in case and int(G.arr(case["gidx"]).max()) >= 2:
        return "offset-recipe out-of-range labels: aggregate_cuda raises instead of corrupting"
    if cid.startswith("dt_uint64_prod_as_int"):
        return "reference result is overflow garbage (int * uint64 promotes to float64, RuntimeWarning: invalid)"
    if cid.endswith("_somenan") and case["func"] == "cumsum":
        return "reference leaks NaN across groups in float cumsum"
    if cid.startswith("dt_float") and "_prod_as_" in cid and ("int" in cid.split("_as_")[1] or "uint" in cid):
        return "per-step truncation of float products into an int dtype"
    return None


def _tolerances(case):
    dt = case.get("dtype", "")
    a_is_f32 = False
    if "a" in case:
        a_is_f32 = G.arr(case["a"]).dtype in (np.float32, np.float16)
    if dt == "float16" or (a_is_f32 and dt == "float16"):
        return dict(rtol=2e-3, atol=1e-3)
    if dt == "float32" or a_is_f32:
        return dict(rtol=1e-5, atol=1e-6)
    return dict(rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("case", G.cases, ids=[c["id"] for c in G.cases])
def test_golden_case(case):
    why = _divergent(case)
    gidx, a, func, kwargs = G.call_args(case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if why and "_sizesmall" in case["id"]:
            with pytest.raises(ValueError):
                aggregate(gidx, a, func=func, **kwargs)
            return
        if why:
            pytest.skip(why)
        if "raises" in case:
            with pytest.raises(G.expected_exception(case)):
                aggregate(gidx, a, func=func, **kwargs)
            return
        res = aggregate(gidx, a, func=func, **kwargs)
    tol = _tolerances(case)
    if case["func"] in ("cumsum", "nancumsum") and case.get("dtype", "").startswith("float"):
        # the reference subtracts group offsets from one global running sum: its error scales with the
        # global magnitude, ours with the group's
        ref = G.arr(case["out"])
        tol["atol"] = max(tol["atol"], tol["rtol"] * float(np.nanmax(np.abs(np.cumsum(np.nan_to_num(G.arr(case["a"]).astype(np.float64).ravel()))))) * 4)
    check_result(G, case, res, **tol)
