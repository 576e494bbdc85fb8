"""ctypes binding of libtmc.so (include/tmc.h).  No torch types cross this boundary.

The library is the product: if it is missing or cannot be loaded this module raises
— there is no CPU fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtmc.so")


class TmcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libtmc error {code}: {msg}")
        self.code = code
        self.msg = msg


# status codes (include/tmc.h)
OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_ZERO_ROW, ERR_NAN, ERR_NEGATIVE, ERR_NOMEM, ERR_COMM, ERR_UNSUPPORTED = (
    0, -1, -2, -3, -4, -5, -6, -7, -8, -9)


class Csr(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
                ("row_ptr", C.c_void_p), ("col_idx", C.c_void_p), ("val", C.c_void_p),
                ("row_offset", C.c_int64), ("n_rows_global", C.c_int64), ("val_type", C.c_int32), ("idx_type", C.c_int32)]


VAL_TYPES = {"float64": 0, "float32": 1, "uint16": 2, "uint8": 3}      # tmc_val_type
IDX_TYPES = {"int32": 0, "uint16": 1}                                   # tmc_idx_type


class Opts(C.Structure):
    _fields_ = [("eigen_group", C.c_int32), ("num_eigen", C.c_int32), ("min_modularity", C.c_double),
                ("min_size", C.c_int32), ("normalize", C.c_int32), ("num_runs", C.c_int32),
                ("seed", C.c_uint64), ("tol", C.c_double), ("max_lanczos", C.c_int32),
                ("max_restarts", C.c_int32), ("max_active", C.c_int32), ("device", C.c_int32),
                ("profile", C.c_int32), ("scatter_mode", C.c_int32), ("sign_stop", C.c_int32), ("warm_start", C.c_int32), ("small_rows", C.c_int32), ("reserved", C.c_int32 * 2)]


class Tree(C.Structure):
    _fields_ = [("n_nodes", C.c_int64), ("n_rows", C.c_int64),
                ("left", C.POINTER(C.c_int64)), ("right", C.POINTER(C.c_int64)), ("parent", C.POINTER(C.c_int64)),
                ("q", C.POINTER(C.c_double)), ("theta", C.POINTER(C.c_double)), ("iters", C.POINTER(C.c_int32)),
                ("item_begin", C.POINTER(C.c_int64)), ("item_end", C.POINTER(C.c_int64)), ("perm", C.POINTER(C.c_int64)),
                ("n_sweeps", C.c_int64), ("spmv_pairs_nnz", C.c_int64), ("kernel_launches", C.c_int64),
                ("engine_ms", C.c_double), ("spmv_ms", C.c_double), ("spmv_alg_bytes", C.c_double),
                ("value_bytes", C.c_int32), ("pad_", C.c_int32), ("p_val", C.POINTER(C.c_double))]


class CsrOut(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_cols", C.c_int64), ("nnz", C.c_int64),
                ("row_ptr", C.POINTER(C.c_int64)), ("col_idx", C.POINTER(C.c_int32)), ("val", C.POINTER(C.c_double)),
                ("kept_rows", C.POINTER(C.c_int64)), ("kept_cols", C.POINTER(C.c_int64)),
                ("n_kept_rows", C.c_int64), ("n_kept_cols", C.c_int64)]


# every symbol include/tmc.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "tmc_opts_default": (None, [C.POINTER(Opts)]),
    "tmc_last_error": (C.c_char_p, []),
    "tmc_version": (C.c_char_p, []),
    "tmc_device_count": (C.c_int, []),
    "tmc_tfidf": (C.c_int, [C.POINTER(Csr), _P, _P]),
    "tmc_row_l2": (C.c_int, [C.POINTER(Csr), _P, _P]),
    "tmc_matrix_upload": (C.c_int, [C.POINTER(Csr), C.c_int, C.c_int, C.POINTER(_P)]),
    "tmc_matrix_adopt_dev": (C.c_int, [C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.POINTER(_P)]),
    "tmc_matrix_free": (None, [_P]),
    "tmc_matrix_info": (C.c_int, [_P, C.POINTER(C.c_int64), C.c_int]),
    "tmc_cache_clear": (None, []),
    "tmc_make_tree": (C.c_int, [C.POINTER(Csr), C.POINTER(Opts), C.POINTER(C.POINTER(Tree))]),
    "tmc_make_tree_dev": (C.c_int, [_P, C.POINTER(Opts), C.POINTER(C.POINTER(Tree))]),
    "tmc_make_tree_dense": (C.c_int, [_P, C.c_int64, C.c_int64, C.POINTER(Opts), C.POINTER(C.POINTER(Tree))]),
    "tmc_tree_free": (None, [C.POINTER(Tree)]),
    "tmc_permutation_test": (C.c_int, [_P, C.POINTER(Tree), C.c_int, C.c_uint64, _P]),
    "tmc_degree": (C.c_int, [_P, _P, C.c_int64, _P]),
    "tmc_second_left": (C.c_int, [_P, _P, C.c_int64, C.POINTER(Opts), _P, C.POINTER(C.c_double), C.POINTER(C.c_int32)]),
    "tmc_modularity": (C.c_int, [_P, _P, C.c_int64, _P, C.POINTER(C.c_double)]),
    "tmc_partition": (C.c_int, [_P, C.c_int64, _P, _P, C.POINTER(C.c_int64)]),
    "tmc_spmv_pair": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "tmc_bench_spmv_pair": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "tmc_mtx_parse": (C.c_int, [C.c_char_p, C.c_int64, C.c_int, C.POINTER(C.POINTER(CsrOut))]),
    "tmc_filter_thresholds": (C.c_int, [C.POINTER(Csr), C.c_double, C.c_double, C.POINTER(C.POINTER(CsrOut))]),
    "tmc_csr_out_free": (None, [C.POINTER(CsrOut)]),
    "tmc_left_singular": (C.c_int, [C.POINTER(Csr), C.c_int, C.c_int, C.POINTER(Opts), _P, _P, C.POINTER(C.c_int)]),
    "tmc_comm_unique_id": (C.c_int, [_P]),
    "tmc_comm_init": (C.c_int, [_P, C.c_int, C.c_int, C.c_int]),
    "tmc_comm_destroy": (None, []),
    "tmc_comm_selftest": (C.c_int, [C.c_int64, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

_lib = None


def load() -> C.CDLL:
    """Load libtmc.so (built in-tree by `__graft_entry__.build()` / `csrc/Makefile`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int):
    if status != 0:
        raise TmcError(status, load().tmc_last_error().decode("utf-8", "replace"))
