"""Loader for the reference's OWN pytest suite (baseline/_ref/numpy_groupies/tests, installed unmodified by
scripts/install_reference.py) with ``aggregate_cuda`` registered as an implementation -- north_star's "passes the
repo's test suite parametrised over all funcs and Forms" and SURVEY section 4.

The reference derives the implementation name from the module name (tests/__init__.py:26-29) and keeps the
registry in module-level lists (tests/__init__.py:16-23, 31-53; test_compare.py:29 TEST_PAIRS), so registering is:
put the module into ``_implementations``, describe what it implements, and (test_compare) add a "cuda/np" pair.
Nothing under baseline/_ref is edited.
"""
import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")

# what aggregate_cuda does not offer (the same role as the reference's table, tests/__init__.py:31-53):
# `rsort` is an alias target nobody implements
NOT_IMPLEMENTED = ("rsort",)


def have_reference():
    return os.path.exists(os.path.join(REF, "numpy_groupies", "tests", "test_generic.py"))


def registry():
    """numpy_groupies.tests with aggregate_cuda as its ONLY implementation (the CPU ones are the reference's business)."""
    if not have_reference():
        pytest.skip("baseline/_ref is missing: run scripts/install_reference.py in the build container", allow_module_level=True)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    rt = importlib.import_module("numpy_groupies.tests")
    from numpy_groupies_b200 import aggregate_cuda
    rt._implementations[:] = [aggregate_cuda]
    rt._implemented_by_impl_name["cuda"] = {"not_implemented": NOT_IMPLEMENTED}
    return rt, aggregate_cuda
