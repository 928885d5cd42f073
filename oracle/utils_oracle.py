"""CPU oracle for the label utilities and ``unpack`` / ``uaggregate``  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy restatements of the helpers that sit either side of ``aggregate``, each citing the reference
lines it follows (paths relative to /root/reference/numpy_groupies).  Parity status: PINNED -- ``oracle/make_golden_utils.py``
runs the real reference functions in the build container and stores their answers in ``tests/golden/utils.npz``;
``tests/test_oracle_golden.py::test_utils_oracle_matches_reference`` replays them through this module bit for bit.
Nothing under ``numpy_groupies_b200/`` imports this file.
"""
import numpy as np


def unpack(group_idx, ret):                               # utils.py:548-553
    return np.asarray(ret)[np.asarray(group_idx)]


def step_count(group_idx):                                # aggregate_numba.py:577-588 (serial loop over value changes)
    g = np.asarray(group_idx)
    if g.size < 1:
        return 0
    return int(np.count_nonzero(g[1:] != g[:-1])) + 1


def step_indices(group_idx):                              # aggregate_numba.py:591-605
    g = np.asarray(group_idx)
    if g.size < 1:
        return np.array([0], dtype=np.int64)              # indices[0] = 0 then indices[-1] = size on a 1-element array
    inner = np.flatnonzero(g[1:] != g[:-1]) + 1
    return np.concatenate(([0], inner, [g.size])).astype(np.int64)


def multi_arange(n):                                      # utils.py:572-594
    n = np.asarray(n)
    if n.ndim != 1:
        raise ValueError("n is supposed to be 1d array.")
    ends = np.cumsum(n)
    total = int(ends[-1])
    starts = ends - n
    seg = np.repeat(np.arange(n.size), n)                 # segment of every output position
    return (np.arange(total) - starts[seg]).astype(int)


def label_contiguous_1d(X):                               # utils.py:597-632
    X = np.asarray(X)
    if X.ndim != 1:
        raise ValueError("this is for 1d masks only.")
    if X.dtype.kind == "b":
        mask = X
        start = np.empty(X.size, dtype=bool)
        start[0] = X[0]
        start[1:] = X[1:] & ~X[:-1]                       # rising edges
    else:
        mask = X != 0
        start = np.empty(X.size, dtype=bool)
        start[0] = X[0] != 0
        start[1:] = (X[1:] != X[:-1]) & mask[1:]          # value changes, zeros never open a block
    labels = np.cumsum(start)
    labels[~mask] = 0
    return labels


def relabel_groups_masked(group_idx, keep_group):         # utils.py:651-681
    group_idx = np.asarray(group_idx)
    keep = np.array(keep_group, dtype=bool)
    keep[0] = True                                        # group 0 cannot be removed
    table = np.zeros(keep.size, dtype=group_idx.dtype)
    table[keep] = np.arange(np.count_nonzero(keep))
    return table[group_idx]


def relabel_groups_unique(group_idx):                     # utils.py:635-648
    group_idx = np.asarray(group_idx)
    keep = np.zeros(int(group_idx.max()) + 1, dtype=bool)
    keep[0] = True
    keep[group_idx] = True
    return relabel_groups_masked(group_idx, keep)
