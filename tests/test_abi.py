"""The C-ABI library builds, loads, and exports every symbol include/aggcuda.h declares (CPU only:
no compute calls)."""
import ctypes
import os
import re

import numpy_groupies_b200._abi as abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "aggcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agc_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    lib = ctypes.CDLL(abi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 9
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in aggcuda.h but not exported"
    assert sorted(abi.EXPORTS) == declared


def test_abi_version_and_struct_sizes():
    lib = abi.load()
    assert lib.agc_abi_version() == abi.ABI_VERSION
    # struct layout must match the C side: 4 + 4 + 8 + 4 + 4 + 64 + 64 + 8
    assert ctypes.sizeof(abi.Index) == 160
    assert ctypes.sizeof(abi.Request) == 16 + 8 + 8 + 8 + 160 + 8 * 11 + 8
    assert ctypes.sizeof(abi.Partial) == 24


def test_argument_errors_are_synchronous_codes():
    lib = abi.load()
    req = abi.Request()
    req.op = 999
    assert lib.agc_workspace_bytes(ctypes.byref(req)) == -1
    assert b"unknown op" in lib.agc_last_error()


def test_no_cpu_fallback_without_cuda():
    import numpy as np
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from numpy_groupies_b200 import aggregate
    with pytest.raises(RuntimeError):
        aggregate(np.array([0, 1, 1]), np.array([1.0, 2.0, 3.0]))


def test_host_label_narrowing_is_pure_host_code():
    """agc_host_narrow_labels (the host half of the H2D path for numpy labels) needs no GPU: int64 / uint64 -> int32,
    everything outside [0, 2^31) becomes -1 (which the kernels report as a bad label, like the original value)."""
    import numpy as np
    lib = abi.load()
    rng = np.random.default_rng(3)
    src = rng.integers(0, 2**31, 100_003)
    src[[5, 77, 1000]] = [-1, 2**31, -2**63]
    dst = np.empty(src.shape[0], dtype=np.int32)
    assert lib.agc_host_narrow_labels(src.ctypes.data, abi.dtype_code(src.dtype), src.shape[0], dst.ctypes.data) == 0
    ref = np.where((src < 0) | (src >= 2**31), -1, src).astype(np.int32)
    np.testing.assert_array_equal(dst, ref)
    u = src.astype(np.uint64)                                # negative values wrap to >= 2^63: still out of range
    assert lib.agc_host_narrow_labels(u.ctypes.data, abi.dtype_code(u.dtype), u.shape[0], dst.ctypes.data) == 0
    np.testing.assert_array_equal(dst, ref)
    i32 = src.astype(np.int32)
    assert lib.agc_host_narrow_labels(i32.ctypes.data, abi.dtype_code(i32.dtype), 4, dst.ctypes.data) != 0
