"""ctypes view of librfb200.so (include/rfb200.h).  Plumbing only: no arithmetic happens in Python.

The library is built in-tree by ``rayforce_b200.build`` (nvcc, sm_100a).  There is no CPU fallback: if the
shared object is missing or no CUDA device is usable every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RFB200_LIB") or os.path.join(HERE, "librfb200.so")   # RFB200_LIB: an alternative build (kernel-geometry sweeps)

# reference type codes (core/rayforce.h:50-62 of the reference)
B8, U8, I16, I32, I64, SYMBOL, DATE, TIME, TIMESTAMP, F64 = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
EQ, NE, LT, GT, LE, GE = range(6)
F_SUM, F_CNT, F_MIN, F_MAX, F_ROWS, F_ALL = 1, 2, 4, 8, 16, 31
ADD, SUB, MUL, DIV, FDIV, MOD, XBAR = range(7)
ROUND, FLOOR, CEIL = range(3)
A_SUM, A_MIN, A_MAX, A_COUNT, A_AVG, A_MED, A_DEV, A_FIRST, A_LAST = range(9)
INDEX_IDS, INDEX_SHIFT = 0, 1
M_AND, M_OR, M_NOT = 0, 1, 2
OK, ERR_TYPE, ERR_LENGTH, ERR_CUDA, ERR_ARG, ERR_NOMEM = 0, -1, -2, -3, -4, -5

NULL_I16 = -(2 ** 15)
NULL_I32 = -(2 ** 31)
NULL_I64 = -(2 ** 63)
INF_I64 = 2 ** 63 - 1

TYPE_SIZE = {B8: 1, U8: 1, I16: 2, I32: 4, DATE: 4, TIME: 4, I64: 8, SYMBOL: 8, TIMESTAMP: 8, F64: 8}


class ScalarValue(C.Union):
    _fields_ = [("i64", C.c_int64), ("f64", C.c_double), ("i32", C.c_int32), ("i16", C.c_int16), ("u8", C.c_uint8)]


class Scalar(C.Structure):
    """rfb_scalar_t"""
    _fields_ = [("type", C.c_int32), ("_pad", C.c_int32), ("v", ScalarValue)]

    @classmethod
    def of(cls, t: int, value) -> "Scalar":
        s = cls()
        s.type = t
        s.v.i64 = 0
        if t == F64:
            s.v.f64 = float(value)
        elif TYPE_SIZE[t] == 8:
            s.v.i64 = int(value)
        elif TYPE_SIZE[t] == 4:
            s.v.i32 = int(value)
        elif TYPE_SIZE[t] == 2:
            s.v.i16 = int(value)
        else:
            s.v.u8 = int(value)
        return s


class Pred(C.Structure):
    """rfb_pred_t"""
    _fields_ = [("op", C.c_int32), ("type", C.c_int32), ("col", C.c_void_p), ("k", Scalar)]


class ColumnFile(C.Structure):
    """rfb_column_file_t"""
    _fields_ = [("type", C.c_int32), ("attrs", C.c_int32), ("len", C.c_int64), ("payload", C.c_void_p), ("map_base", C.c_void_p),
                ("map_bytes", C.c_size_t)]


class Fold(C.Structure):
    """rfb_fold_t"""
    _fields_ = [("rows", C.c_int64), ("nonnull", C.c_int64), ("sum_i64", C.c_int64), ("sum_f64", C.c_double),
                ("min_i64", C.c_int64), ("max_i64", C.c_int64), ("min_f64", C.c_double), ("max_f64", C.c_double),
                ("sum_f64_err", C.c_double)]


class GroupInfo(C.Structure):
    """rfb_group_info_t"""
    _fields_ = [("index_type", C.c_int32), ("dense", C.c_int32), ("groups", C.c_int64), ("min", C.c_int64),
                ("max", C.c_int64), ("range", C.c_int64)]


class RfbError(RuntimeError):
    def __init__(self, code: int, text: str):
        super().__init__("%s (rfb_status %d)" % (text, code))
        self.code = code
        # the reference renders these as the Rayfall errors "type" / "length" (core/error.h:86-97)
        self.kind = {ERR_TYPE: "type", ERR_LENGTH: "length", ERR_CUDA: "cuda", ERR_ARG: "arg", ERR_NOMEM: "limit"}.get(code, "?")


_vp, _i64, _ci, _u64, _sz = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_size_t
_P = C.POINTER

# name -> (restype, argtypes).  Must list every function include/rfb200.h declares (tests/test_abi.py checks).
SIGNATURES = {
    "rfb_abi_version": (_ci, []),
    "rfb_last_error": (C.c_char_p, []),
    "rfb_device_count": (_ci, []),
    "rfb_ctx_create": (_ci, [_ci, _P(_vp)]),
    "rfb_ctx_destroy": (None, [_vp]),
    "rfb_ctx_set_stream": (_ci, [_vp, _vp]),
    "rfb_ctx_set_result_ptr": (_ci, [_vp, _vp]),
    "rfb_ctx_stream": (_vp, [_vp]),
    "rfb_ctx_sm_count": (_ci, [_vp]),
    "rfb_sync": (_ci, [_vp]),
    "rfb_launch_count": (_i64, [_vp]),
    "rfb_dev_alloc": (_ci, [_vp, _sz, _P(_vp)]),
    "rfb_dev_free": (_ci, [_vp, _vp]),
    "rfb_dev_memset": (_ci, [_vp, _vp, _ci, _sz]),
    "rfb_dev_mem_info": (_ci, [_vp, C.POINTER(_sz), C.POINTER(_sz)]),
    "rfb_host_pin": (_ci, [_vp, _sz]),
    "rfb_host_unpin": (_ci, [_vp]),
    "rfb_host_alloc_pinned": (_ci, [_sz, _P(_vp)]),
    "rfb_host_free_pinned": (_ci, [_vp]),
    "rfb_h2d": (_ci, [_vp, _vp, _vp, _sz]),
    "rfb_d2h": (_ci, [_vp, _vp, _vp, _sz]),
    "rfb_d2h_sync_plain": (_ci, [_vp, _vp, _vp, _sz]),
    "rfb_fill_splitmix_dev": (_ci, [_vp, _ci, _vp, _i64, _u64, _u64, _i64, _i64, C.c_double]),
    "rfb_fold_dev": (_ci, [_vp, _ci, _ci, _vp, _i64, _P(Fold)]),
    "rfb_fold_result": (_ci, [_vp, _P(Fold)]),
    "rfb_filter_fold_dev": (_ci, [_vp, _ci, _ci, _vp, _P(Scalar), _ci, _ci, _vp, _i64, _P(Fold)]),
    "rfb_multi_filter_fold_dev": (_ci, [_vp, _ci, _P(Pred), _ci, _ci, _ci, _vp, _i64, _P(Fold)]),
    "rfb_fma_fold_dev": (_ci, [_vp, _ci, _vp, _vp, _vp, _i64, _P(Fold)]),
    "rfb_cmp_dev": (_ci, [_vp, _ci, _ci, _vp, _i64, _P(Scalar), _ci, _vp, _i64, _P(Scalar), _vp]),
    "rfb_mask_logic_dev": (_ci, [_vp, _ci, _vp, _i64, _vp, _i64, C.c_uint8, _vp]),
    "rfb_where_dev": (_ci, [_vp, _vp, _i64, _vp, _P(_i64)]),
    "rfb_cmp_where_dev": (_ci, [_vp, _ci, _ci, _vp, _i64, _P(Scalar), _vp, _P(_i64)]),
    "rfb_gather_dev": (_ci, [_vp, _ci, _vp, _vp, _i64, _vp]),
    "rfb_gather_fold_dev": (_ci, [_vp, _ci, _ci, _vp, _vp, _i64, _P(Fold)]),
    "rfb_binop_type": (_ci, [_ci, _ci, _ci]),
    "rfb_binop_type_form": (_ci, [_ci, _ci, _ci, _ci]),
    "rfb_binop_dev": (_ci, [_vp, _ci, _ci, _vp, _i64, _P(Scalar), _ci, _vp, _i64, _P(Scalar), _vp]),
    "rfb_unop_f64_dev": (_ci, [_vp, _ci, _vp, _i64, _vp]),
    "rfb_group_i64_dev": (_ci, [_vp, _vp, _vp, _i64, _vp, _vp, _P(GroupInfo)]),
    "rfb_group_keys_i64_dev": (_ci, [_vp, _ci, _P(_vp), _vp, _i64, _vp, _vp, _P(GroupInfo)]),
    "rfb_aggr_type": (_ci, [_ci, _ci]),
    "rfb_aggr_dev": (_ci, [_vp, _ci, _ci, _vp, _vp, _vp, _i64, _i64, _vp]),
    "rfb_aggr_last_dev": (_ci, [_vp, _ci, _vp, _vp, _vp, _i64, _i64, _i64, _vp]),
    "rfb_group_sum_count_host": (_ci, [_vp, _ci, _vp, _vp, _i64, _ci, _ci, _vp, _P(Scalar), _i64, _vp, _vp, _vp, _P(_i64), _P(_i64)]),
    "rfb_fma_fold_host": (_ci, [_vp, _ci, _vp, _vp, _vp, _i64, _P(Fold), _P(_i64)]),
    "rfb_window_aggr_dev": (_ci, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _ci, _ci, _ci, _vp, _vp]),
    "rfb_options_reload": (None, []),
    "rfb_peer_mailbox_create": (_ci, [_vp, _vp]),
    "rfb_peer_mailbox_bind": (_ci, [_vp, _ci, _ci, _vp]),
    "rfb_fold_allreduce_peers": (_ci, [_vp, _ci, _P(Fold)]),
    "rfb_fold_peers_result": (_ci, [_vp, _P(Fold)]),
    "rfb_peer_groups_create": (_ci, [_vp, _i64, _vp]),
    "rfb_peer_groups_bind": (_ci, [_vp, _ci, _ci, _vp]),
    "rfb_group_merge_peers": (_ci, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _P(_i64)]),
    "rfb_mgpu_create": (_ci, [_ci, _P(_vp)]),
    "rfb_mgpu_destroy": (None, [_vp]),
    "rfb_mgpu_devices": (_ci, [_vp]),
    "rfb_mgpu_ctx": (_vp, [_vp, _ci]),
    "rfb_mgpu_filter_fold_host": (_ci, [_vp, _ci, _ci, _vp, _P(Scalar), _ci, _ci, _vp, _i64, _i64, _P(Fold), _P(_i64)]),
    "rfb_mgpu_group_sum_count_host": (_ci, [_vp, _ci, _vp, _vp, _i64, _ci, _ci, _vp, _P(Scalar), _i64, _vp, _vp, _vp, _P(_i64), _P(_i64)]),
    "rfb_group_rows_dev": (_ci, [_vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "rfb_med_dev": (_ci, [_vp, _ci, _vp, _i64, _P(C.c_double)]),
    "rfb_stddev_dev": (_ci, [_vp, _ci, _vp, _i64, _P(C.c_double)]),
    "rfb_find_rows_dev": (_ci, [_vp, _ci, _P(_vp), _i64, _P(_vp), _i64, _vp]),
    "rfb_inner_join_dev": (_ci, [_vp, _ci, _P(_vp), _i64, _P(_vp), _i64, _vp, _vp, _P(_i64)]),
    "rfb_asof_join_dev": (_ci, [_vp, _ci, _P(_vp), _ci, _vp, _i64, _P(_vp), _vp, _i64, _vp]),
    "rfb_distinct_i64_dev": (_ci, [_vp, _vp, _i64, _vp, _P(_i64)]),
    "rfb_window_join_dev": (_ci, [_vp, _ci, _P(_vp), _vp, _i64, _P(_vp), _i64, _vp, _vp, _ci, _ci, _ci, _vp, _vp]),
    "rfb_group_sum_count_dev": (_ci, [_vp, _ci, _vp, _vp, _i64, _ci, _ci, _vp, _P(Scalar), _i64, _vp, _vp, _vp, _P(_i64)]),
    "rfb_sort_dev": (_ci, [_vp, _ci, _vp, _i64, _ci, _vp]),
    "rfb_column_file_open": (_ci, [C.c_char_p, _P(ColumnFile)]),
    "rfb_column_file_close": (_ci, [_P(ColumnFile)]),
    "rfb_column_file_write": (_ci, [C.c_char_p, _ci, _ci, _vp, _i64]),
    "rfb_filter_fold_host": (_ci, [_vp, _ci, _ci, _vp, _P(Scalar), _ci, _ci, _vp, _i64, _i64, _P(Fold), _P(_i64)]),
    "rfb_fold_host": (_ci, [_vp, _ci, _ci, _vp, _i64, _i64, _P(Fold), _P(_i64)]),
}

_lib = None


def load(path: str = LIB_PATH) -> C.CDLL:
    """dlopen librfb200.so and attach signatures.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise RfbError(ERR_CUDA, "librfb200.so is not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`; "
                                 "there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name, None)
        if fn is None:
            continue  # reported by tests/test_abi.py; calling it raises AttributeError
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != OK:
        raise RfbError(rc, (load().rfb_last_error() or b"").decode(errors="replace"))
