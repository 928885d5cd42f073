"""Loader for tests/golden/utils.npz (written by oracle/make_golden_utils.py from the real reference)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def cases():
    z = np.load(os.path.join(HERE, "golden", "utils.npz"))
    out = {}
    for key in z.files:
        cid, kind, what = key.split("__")
        out.setdefault((int(cid[1:]), kind), {})[what] = z[key]
    return [(cid, kind, d) for (cid, kind), d in sorted(out.items())]


def run(kind, d, impl, aggregate=None):
    """Apply the implementation ``impl`` (a module / namespace with the reference's function names) to one case."""
    x = d["in_x"]
    if kind == "step_count":
        return np.asarray(impl.step_count(x))
    if kind == "relabel_groups_masked":
        return impl.relabel_groups_masked(x, d["in_keep"].copy())
    if kind == "unpack":
        return impl.unpack(x, d["in_ret"])
    if kind.startswith("uaggregate_"):
        return impl.unpack(x, aggregate(x, d["in_a"], func=kind.split("_", 1)[1]))
    return getattr(impl, kind)(x)
