"""Loader for tests/golden (written by oracle/make_golden.py from the real reference)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

CALLABLES = {
    "len_gt1": lambda x: len(x) > 1,
    "sum_method": lambda x: x.sum(),
    "ptp": lambda x: x.max() - x.min(),
    "median": np.median,
}
_BUILTIN_BASES = (NotImplementedError, OverflowError, TypeError, ValueError, IndexError, ZeroDivisionError)


def dec(v):
    if isinstance(v, dict):
        if "f" in v:
            return float(v["f"])
        if "t" in v:
            return tuple(dec(x) for x in v["t"])
        if "dtype" in v:
            return np.dtype(v["dtype"])
    return v


class Golden:
    def __init__(self):
        with open(os.path.join(HERE, "golden", "manifest.json")) as f:
            self.manifest = json.load(f)
        self._npz = np.load(os.path.join(HERE, "golden", "cases.npz"))
        self._cache = {}
        self.cases = self.manifest["cases"]
        self.by_id = {c["id"]: c for c in self.cases}

    def arr(self, key):
        if key not in self._cache:
            self._cache[key] = self._npz[key]
        return self._cache[key].copy()

    def call_args(self, case):
        gidx = self.arr(case["gidx"])
        a = self.arr(case["a"]) if "a" in case else dec(case["a_scalar"])
        func = case["func"] if isinstance(case["func"], str) else CALLABLES[case["func"]["callable"]]
        kwargs = {k: dec(v) for k, v in case["kwargs"].items()}
        if "dtype" in kwargs and isinstance(kwargs["dtype"], str):
            kwargs["dtype"] = np.dtype(kwargs["dtype"])
        if isinstance(kwargs.get("size"), tuple):
            kwargs["size"] = list(kwargs["size"])
        return gidx, a, func, kwargs

    @staticmethod
    def expected_exception(case):
        """Reference exception class, widened to its builtin base (numpy's UFuncTypeError is a TypeError)."""
        name = case["raises"]
        if name in ("UFuncTypeError", "_UFuncOutputCastingError", "_UFuncNoLoopError"):
            return TypeError
        return {b.__name__: b for b in _BUILTIN_BASES}[name]


def check_result(golden, case, res, rtol=1e-12, atol=0.0, exact_float=False):
    """Compare ``res`` with the reference's recorded answer: dtype, shape, layout, values."""
    if "object_lens" in case:
        lens = golden.arr(case["object_lens"]); vals = golden.arr(case["object_vals"])
        assert res.dtype == object and len(res) == len(lens)
        pos = 0
        for item, n in zip(res, lens):
            if n < 0:
                assert np.ndim(item) == 0
            else:
                np.testing.assert_array_equal(np.asarray(item), vals[pos:pos + n]); pos += n
        return
    ref = golden.arr(case["out"])
    assert isinstance(res, np.ndarray), type(res)
    assert res.dtype.name == case["dtype"], f"dtype {res.dtype.name} != reference {case['dtype']}"
    assert list(res.shape) == case["shape"], f"shape {res.shape} != reference {case['shape']}"
    if not (isinstance(case["func"], str) and "arg" in case["func"]):
        # arg* on Form 3: the reference reshapes a strided view of np.unravel_index's output, which is
        # neither C- nor F-contiguous; only its values and shape are contractual there
        assert bool(np.isfortran(res)) == case["fortran"]
    if ref.dtype.kind in "fc" and not exact_float:
        np.testing.assert_allclose(res, ref, rtol=rtol, atol=atol, equal_nan=True)
    else:
        np.testing.assert_array_equal(res, ref)
