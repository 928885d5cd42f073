"""The reference's tests/test_generic.py, run against aggregate_cuda (see tests/refsuite.py)."""
import pytest

import refsuite

pytestmark = pytest.mark.gpu
_rt, _cuda = refsuite.registry()
from numpy_groupies.tests.test_generic import *  # noqa: E402,F401,F403  (test functions + the aggregate_all fixture)
