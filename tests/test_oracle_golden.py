"""Pins oracle/aggregate_oracle.py against the real reference's recorded answers (CPU only)."""
import warnings

import numpy as np
import pytest

from golden_io import Golden, check_result
from oracle import aggregate_oracle as oracle

G = Golden()


@pytest.mark.parametrize("case", G.cases, ids=[c["id"] for c in G.cases])
def test_oracle_matches_reference(case):
    gidx, a, func, kwargs = G.call_args(case)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if "raises" in case:
            with pytest.raises(G.expected_exception(case)):
                oracle.aggregate(gidx, a, func=func, **kwargs)
            return
        res = oracle.aggregate(gidx, a, func=func, **kwargs)
    # same primitives in the same order => bit-identical floats too
    check_result(G, case, res, exact_float=True)


def test_utils_oracle_matches_reference():
    """oracle/utils_oracle.py == the reference's label utilities / unpack on every stored case, bit for bit."""
    import utils_golden
    from oracle import aggregate_oracle, utils_oracle
    n = 0
    for cid, kind, d in utils_golden.cases():
        got = utils_golden.run(kind, d, utils_oracle, aggregate_oracle.aggregate)
        ref = d["out"]
        assert np.asarray(got).dtype == ref.dtype, (cid, kind, np.asarray(got).dtype, ref.dtype)
        np.testing.assert_array_equal(got, ref, err_msg=f"case {cid} {kind}")
        n += 1
    assert n >= 30
