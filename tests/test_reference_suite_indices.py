"""The reference's tests/test_indices.py (step_count / step_indices) against aggregate_cuda's device versions."""
import sys

import pytest

import refsuite

pytestmark = pytest.mark.gpu
_rt, _cuda = refsuite.registry()
if not hasattr(_cuda, "step_indices"):
    pytest.skip("aggregate_cuda.step_indices is not built yet", allow_module_level=True)
_saved = _rt.aggregate_numba
_rt.aggregate_numba = _cuda                      # test_indices.py:4-6 builds its implementation list from this name
sys.modules.pop("numpy_groupies.tests.test_indices", None)
try:
    from numpy_groupies.tests.test_indices import *  # noqa: E402,F401,F403
finally:
    _rt.aggregate_numba = _saved
