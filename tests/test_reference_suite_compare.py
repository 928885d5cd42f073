"""The reference's tests/test_compare.py with the pair "cuda/np": aggregate_cuda against aggregate_numpy at the
reference's own tolerance (rtol 1e-10, dtype equality; test_compare.py:102-144)."""
import pytest

import refsuite

pytestmark = pytest.mark.gpu
_rt, _cuda = refsuite.registry()
import numpy_groupies.tests.test_compare as _tc  # noqa: E402
from numpy_groupies.tests.test_compare import *  # noqa: E402,F401,F403

# the reference's fixture body, called for the "numba/np" pair (same data recipe, group_cnt = 1000, aggregate_numpy as
# func_ref), then pointed at aggregate_cuda
_orig = getattr(_tc.aggregate_cmp, "__wrapped__", None) or _tc.aggregate_cmp.__pytest_wrapped__.obj


class _Req:
    param = "numba/np"


@pytest.fixture(params=["cuda/np"], scope="module")
def aggregate_cmp(request):
    d = _orig(_Req())
    d["impl"], d["func"], d["name"], d["test_pair"] = _cuda, _cuda.aggregate, "cuda", "cuda/np"
    return d
