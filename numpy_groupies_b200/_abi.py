"""ctypes binding of libaggcuda.so (include/aggcuda.h).  No fallback: if the shared library is
missing or was built for another ABI this module raises, and so does everything above it."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libaggcuda.so")
ABI_VERSION = 5
MAX_NDIM = 8

# ---- enums (mirror include/aggcuda.h) ----------------------------------------------------------
DT = {"bool": 0, "int8": 1, "uint8": 2, "int16": 3, "uint16": 4, "int32": 5, "uint32": 6, "int64": 7,
      "uint64": 8, "float32": 9, "float64": 10}
DT_NAME = {v: k for k, v in DT.items()}
OP = {"sum": 0, "prod": 1, "len": 2, "last": 3, "first": 4, "all": 5, "any": 6, "min": 7, "max": 8, "argmax": 9,
      "argmin": 10, "mean": 11, "sumofsquares": 12, "var": 13, "std": 13, "allnan": 14, "anynan": 15,
      "cumsum": 16, "cumprod": 17, "cummax": 18, "cummin": 19, "sort": 20, "minmax": 21}
F_NANSKIP, F_SQRT, F_REVERSE, F_ASSUME_SORTED, F_NO_FAST = 0x1, 0x2, 0x4, 0x8, 0x10
F_NEED_TOUCHED, F_ACC_ONLY, F_FIN_ONLY, F_FILL_NAN, F_NO_INIT, F_PASS2 = 0x20, 0x40, 0x80, 0x100, 0x200, 0x400
F_TOUCHED_BY_COUNT, F_KEYS_IN_T2, F_DETERMINISTIC = 0x800, 0x1000, 0x2000
LB_STEPS, LB_STEP_COUNT, LB_CONTIG, LB_KEEP, LB_CUMSUM = 0, 1, 2, 3, 4
FORM_FLAT, FORM_AXIS, FORM_MULTI = 0, 1, 2
ST_BADLABEL, ST_PRECISION, ST_UNSORTED, ST_REGIME = 0, 1, 2, 3
REDUCE_NAMES = {0: "sum", 1: "max", 2: "min", 3: "prod"}


class Index(C.Structure):
    _fields_ = [("form", C.c_int32), ("dtype", C.c_int32), ("labels", C.c_void_p), ("ndim", C.c_int32),
                ("axis", C.c_int32), ("dims", C.c_int64 * MAX_NDIM), ("strides", C.c_int64 * MAX_NDIM),
                ("axis_size", C.c_int64)]


class Request(C.Structure):
    _fields_ = [("op", C.c_int32), ("flags", C.c_int32), ("a_dtype", C.c_int32), ("out_dtype", C.c_int32),
                ("a", C.c_void_p), ("a_scalar_f64", C.c_double), ("a_scalar_i64", C.c_int64), ("index", Index),
                ("n", C.c_int64), ("size", C.c_int64), ("out", C.c_void_p), ("fill_f64", C.c_double),
                ("fill_i64", C.c_int64), ("ddof", C.c_double), ("index_base", C.c_int64),
                ("arg_stride", C.c_int64), ("arg_len", C.c_int64), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_int64), ("status", C.c_void_p)]


class Partial(C.Structure):
    _fields_ = [("offset_bytes", C.c_int64), ("count", C.c_int64), ("dtype", C.c_int32), ("reduce", C.c_int32)]


class PackSeg(C.Structure):
    _fields_ = [("table", C.c_void_p), ("count", C.c_int64), ("dtype", C.c_int32), ("xform", C.c_int32)]


PK_COPY, PK_NEG, PK_FLAG, PK_NEGFLAG = 0, 1, 2, 3

EXPORTS = ("agc_abi_version", "agc_last_error", "agc_workspace_bytes", "agc_partials", "agc_aggregate",
           "agc_label_minmax", "agc_unpack", "agc_group_order_workspace_bytes", "agc_group_order",
           "agc_profile_enable", "agc_profile_last_ms", "agc_profile_ms", "agc_launch_count", "agc_tuning",
           "agc_label_scan_workspace_bytes", "agc_label_scan", "agc_mark_present", "agc_multi_arange",
           "agc_pack_tables", "agc_rank_fold", "agc_apply_carry", "agc_owner_pack", "agc_ordered_accumulate",
           "agc_host_narrow_labels", "agc_host_copy_stream", "agc_convert_f16")

_lib = None


def load():
    """Load libaggcuda.so (once).  Raises RuntimeError if it is absent -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  aggregate_cuda has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise RuntimeError(f"{LIB_PATH} lacks symbols {missing}")
    lib.agc_abi_version.restype = C.c_int
    lib.agc_last_error.restype = C.c_char_p
    lib.agc_workspace_bytes.restype = C.c_int64
    lib.agc_workspace_bytes.argtypes = [C.POINTER(Request)]
    lib.agc_partials.restype = C.c_int
    lib.agc_partials.argtypes = [C.POINTER(Request), C.POINTER(Partial), C.c_int]
    lib.agc_aggregate.restype = C.c_int
    lib.agc_aggregate.argtypes = [C.POINTER(Request), C.c_void_p]
    lib.agc_label_minmax.restype = C.c_int
    lib.agc_label_minmax.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]
    lib.agc_unpack.restype = C.c_int
    lib.agc_unpack.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p,
                               C.c_void_p, C.c_void_p]
    lib.agc_group_order_workspace_bytes.restype = C.c_int64
    lib.agc_group_order_workspace_bytes.argtypes = [C.c_int64, C.c_int64]
    lib.agc_group_order.restype = C.c_int
    lib.agc_group_order.argtypes = [C.POINTER(Index), C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.c_void_p, C.c_void_p]
    lib.agc_label_scan_workspace_bytes.restype = C.c_int64
    lib.agc_label_scan_workspace_bytes.argtypes = [C.c_int64]
    lib.agc_label_scan.restype = C.c_int
    lib.agc_label_scan.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_int32, C.c_void_p,
                                   C.c_void_p, C.c_int64, C.c_void_p]
    lib.agc_mark_present.restype = C.c_int
    lib.agc_mark_present.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.agc_multi_arange.restype = C.c_int
    lib.agc_multi_arange.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.agc_pack_tables.restype = C.c_int
    lib.agc_pack_tables.argtypes = [C.POINTER(PackSeg), C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    lib.agc_ordered_accumulate.restype = C.c_int
    lib.agc_ordered_accumulate.argtypes = [C.POINTER(Request), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.agc_owner_pack.restype = C.c_int
    lib.agc_owner_pack.argtypes = [C.POINTER(Request), C.c_void_p, C.c_int32, C.c_void_p]
    lib.agc_rank_fold.restype = C.c_int
    lib.agc_rank_fold.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
    lib.agc_apply_carry.restype = C.c_int
    lib.agc_apply_carry.argtypes = [C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_void_p, C.c_int32, C.c_int32,
                                    C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
    lib.agc_host_narrow_labels.restype = C.c_int
    lib.agc_host_narrow_labels.argtypes = [C.c_void_p, C.c_int32, C.c_int64, C.c_void_p]
    lib.agc_host_copy_stream.restype = C.c_int
    lib.agc_host_copy_stream.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
    lib.agc_convert_f16.restype = C.c_int
    lib.agc_convert_f16.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    lib.agc_profile_enable.restype = C.c_int
    lib.agc_profile_enable.argtypes = [C.c_int]
    lib.agc_profile_last_ms.restype = C.c_float
    lib.agc_profile_ms.restype = C.c_float
    lib.agc_profile_ms.argtypes = [C.c_int]
    lib.agc_launch_count.restype = C.c_int64
    lib.agc_tuning.restype = C.c_int
    lib.agc_tuning.argtypes = [C.c_int, C.c_int]
    if lib.agc_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libaggcuda ABI {lib.agc_abi_version()} != expected {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def last_error():
    return load().agc_last_error().decode()


_DT_CACHE = {}


def dtype_code(np_dtype):
    # np.dtype(...).name is surprisingly slow (string assembly on every access): cache by the dtype object itself
    try:
        return _DT_CACHE[np_dtype]
    except (KeyError, TypeError):
        code = DT[np.dtype(np_dtype).name]
        try:
            _DT_CACHE[np_dtype] = code
        except TypeError:
            pass
        return code
