#!/usr/bin/env python
"""Freeze the REAL reference's answers for the label utilities into tests/golden/utils.npz.

    python oracle/make_golden_utils.py        # needs /root/reference (or baseline/_ref)

Cases: the docstring examples of utils.py:572-681, the vectors of the reference's tests/test_indices.py:17-29 and
tests/test_utils.py:12-25, plus seeded random inputs (including multi-tile sizes for the device scans).
Test infrastructure only.
"""
import os
import sys

import numpy as np

REF_ROOT = os.environ.get("NPG_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "utils.npz")


def main():
    sys.path.insert(0, REF_ROOT)
    import numpy_groupies as npg
    from numpy_groupies import utils
    from numpy_groupies import aggregate_numba as nb

    rng = np.random.RandomState(100)
    store = {}
    k = 0

    def add(kind, out, **inputs):
        nonlocal k
        for name, v in inputs.items():
            store[f"c{k}__{kind}__in_{name}"] = np.asarray(v)
        store[f"c{k}__{kind}__out"] = np.asarray(out)
        k += 1

    # step_count / step_indices: tests/test_indices.py vectors + random runs
    steps_inputs = [np.array([1, 1, 1, 2, 2, 3, 3, 4, 5, 2, 2]), np.array([1, 1, 1, 2, 2, 3, 3, 4, 4, 2, 2]), np.array([7]),
                    np.repeat(rng.randint(0, 50, 3000), rng.randint(1, 9, 3000)), rng.randint(0, 3, 10000),
                    np.zeros(5000, dtype=np.int32), np.arange(4100, dtype=np.int64)]
    for g in steps_inputs:
        add("step_count", nb.step_count(g), x=g)
        add("step_indices", nb.step_indices(g), x=g)
    # multi_arange
    for n in (np.array([0, 0, 3, 0, 0, 2, 0, 2, 1]), rng.randint(0, 6, 5000), np.array([5]), rng.randint(0, 3, 100).astype(np.int32)):
        add("multi_arange", utils.multi_arange(n), x=n)
    # label_contiguous_1d
    for X in (np.array([0, 1, 1, 0, 0, 1, 0, 0, 0, 1, 1, 1], dtype=bool), np.array([0, 3, 3, 0, 0, 5, 5, 5, 1, 1, 0, 2]),
              rng.rand(6000) < 0.4, np.repeat(rng.randint(0, 4, 2500), rng.randint(1, 5, 2500)), np.ones(3000, dtype=bool),
              np.array([2, 2, 0, 2], dtype=np.int8)):
        add("label_contiguous_1d", utils.label_contiguous_1d(X), x=X)
    # relabel_groups_masked / unique
    g = np.array([0, 3, 3, 3, 0, 2, 5, 2, 0, 1, 1, 0, 3, 5, 5])
    add("relabel_groups_masked", utils.relabel_groups_masked(g, np.array([0, 1, 0, 1, 1, 1], dtype=bool)), x=g, keep=np.array([0, 1, 0, 1, 1, 1], dtype=bool))
    add("relabel_groups_unique", utils.relabel_groups_unique(g), x=g)
    g2 = rng.randint(0, 4000, 20000) * 3
    keep2 = rng.rand(int(g2.max()) + 1) < 0.5
    add("relabel_groups_masked", utils.relabel_groups_masked(g2, keep2.copy()), x=g2, keep=keep2)
    add("relabel_groups_unique", utils.relabel_groups_unique(g2), x=g2)
    g3 = rng.randint(0, 100, 5000).astype(np.int32)
    add("relabel_groups_unique", utils.relabel_groups_unique(g3), x=g3)
    # unpack / uaggregate (tests/test_utils.py:12-25 recipes)
    gi = np.repeat(rng.permutation(10), 3)
    vals = rng.randn(10)
    add("unpack", utils.unpack(gi, vals), x=gi, ret=vals)
    gi = np.repeat(np.arange(3000), 10)
    vals = rng.randn(3000)
    add("unpack", utils.unpack(gi, vals), x=gi, ret=vals)
    gi = rng.randint(0, 300, 3000)
    a = rng.rand(3000)
    for func in ("sum", "max", "mean", "len"):
        add("uaggregate_" + func, utils.unpack(gi, npg.aggregate_np(gi, a, func=func)), x=gi, a=a)
    np.savez_compressed(OUT, **store)
    print(f"{k} cases -> {OUT} ({os.path.getsize(OUT)} bytes)")


if __name__ == "__main__":
    main()
