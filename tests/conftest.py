"""pytest configuration: registers the ``gpu`` marker and puts the repo root on sys.path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "deselect_if(func): drop parametrisations for which func(**params) is true")


def pytest_collection_modifyitems(config, items):
    """The reference's suite (run from tests/test_reference_suite_*.py) marks invalid parameter combinations with
    ``deselect_if``; its own root conftest.py:7-26 honours the marker, ours has to as well."""
    keep, drop = [], []
    for item in items:
        mark = item.get_closest_marker("deselect_if")
        callspec = getattr(item, "callspec", None)
        if mark is not None and callspec is not None and mark.kwargs["func"](**callspec.params):
            drop.append(item)
        else:
            keep.append(item)
    if drop:
        config.hook.pytest_deselected(items=drop)
        items[:] = keep
