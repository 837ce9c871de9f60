"""Python view of the reference-facing operator layer (include/rfb200_ops.h, librfb200_ops.so) with its builtin
malloc host.  Objects are the reference's obj_t layout; this module only builds them from numpy arrays, calls the
operators (same names as the reference's, without the `rfb_` prefix) and reads results back.  No arithmetic here."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import RfbError

HERE = os.path.dirname(os.path.abspath(__file__))
OPS_PATH = os.path.join(HERE, "librfb200_ops.so")

LIST, MAPFILTER, MAPGROUP, T_NULL, T_ERR = 0, 71, 72, 126, 127
NP_OF = {capi.B8: np.uint8, capi.U8: np.uint8, capi.I16: np.int16, capi.I32: np.int32, capi.DATE: np.int32,
         capi.TIME: np.int32, capi.I64: np.int64, capi.SYMBOL: np.int64, capi.TIMESTAMP: np.int64, capi.F64: np.float64}

UNARY = ["ray_where", "ray_sum", "ray_min", "ray_max", "ray_avg", "ray_cnt", "ray_round", "ray_floor", "ray_ceil",
         "ray_sort_asc", "ray_sort_desc", "ray_med", "ray_dev", "ray_distinct", "ray_not", "ray_asc", "ray_desc"]
BINARY = ["ray_eq", "ray_ne", "ray_lt", "ray_gt", "ray_le", "ray_ge", "filter_map", "filter_collect", "ray_add", "ray_sub",
          "ray_mul", "ray_div", "ray_fdiv", "ray_mod", "ray_xbar", "index_group", "group_map", "aggr_sum", "aggr_min", "aggr_max",
          "aggr_count", "aggr_avg", "aggr_med", "aggr_stddev", "aggr_row", "aggr_collect", "ray_find", "ray_in", "where_lt_sum",
          "aggr_first", "aggr_last", "index_group_list", "ray_xasc", "ray_xdesc"]
TERNARY_I64 = ["index_left_join_obj", "index_inner_join_obj"]      # (obj, obj, int64 len)


class HostApi(C.Structure):
    _fields_ = [("vector", C.CFUNCTYPE(C.c_void_p, C.c_int8, C.c_int64)), ("atom", C.CFUNCTYPE(C.c_void_p, C.c_int8)),
                ("clone_obj", C.CFUNCTYPE(C.c_void_p, C.c_void_p)), ("drop_obj", C.CFUNCTYPE(None, C.c_void_p)),
                ("err_type", C.CFUNCTYPE(C.c_void_p)), ("err_length", C.CFUNCTYPE(C.c_void_p)),
                ("err_limit", C.CFUNCTYPE(C.c_void_p)), ("null_obj", C.c_void_p), ("executors", C.c_void_p)]


class OpsError(Exception):
    """the operator returned the host's ERR_OBJ; `.kind` is the reference's error name ("type", "length", "limit")"""

    def __init__(self, kind):
        super().__init__(kind)
        self.kind = kind


class Declined(Exception):
    """the GPU layer declined the operand (NULL): the reference's CPU body would run instead"""


class Ops:
    _inst = None

    @classmethod
    def get(cls, device: int = 0) -> "Ops":
        if cls._inst is None:
            cls._inst = cls(device)
        return cls._inst

    def __init__(self, device: int = 0):
        capi.load()      # librfb200.so first (the ops library links against it)
        if not os.path.exists(OPS_PATH):
            raise RfbError(capi.ERR_CUDA, "librfb200_ops.so is not built")
        L = self.L = C.CDLL(OPS_PATH)
        L.rfb_ops_builtin_host.restype = C.POINTER(HostApi)
        L.rfb_ops_init.argtypes = [C.POINTER(HostApi), C.c_int]
        L.rfb_ops_last_error.restype = C.c_char_p
        L.rfb_ops_builtin_last_error.restype = C.c_char_p
        L.rfb_ops_launches.restype = C.c_int64
        L.rfb_ops_set_min_rows.argtypes = [C.c_int64]
        self.host = L.rfb_ops_builtin_host().contents
        rc = L.rfb_ops_init(L.rfb_ops_builtin_host(), device)
        if rc != 0:
            raise RfbError(rc, (L.rfb_ops_last_error() or b"").decode())
        for n in UNARY:
            f = getattr(L, "rfb_" + n)
            f.restype, f.argtypes = C.c_void_p, [C.c_void_p]
        for n in BINARY:
            f = getattr(L, "rfb_" + n)
            f.restype, f.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p]
        for n in TERNARY_I64:
            f = getattr(L, "rfb_" + n)
            f.restype, f.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int64]
        L.rfb_index_asof_join_obj.restype, L.rfb_index_asof_join_obj.argtypes = C.c_void_p, [C.c_void_p] * 4
        L.rfb_where_fold.restype, L.rfb_where_fold.argtypes = C.c_void_p, [C.POINTER(C.c_void_p), C.c_int64]
        L.rfb_at_ids.restype, L.rfb_at_ids.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int64]
        L.rfb_mask_logic_inplace.restype, L.rfb_mask_logic_inplace.argtypes = C.c_int, [C.c_int, C.c_void_p, C.c_void_p]
        L.rfb_ops_set_gate.argtypes = [C.c_int]
        L.rfb_ops_set_gate(0)     # tests call single operators on cold vectors: no cost gate
        self.NULL = self.host.null_obj

    # ---- objects
    def vec(self, t, arr):
        arr = np.ascontiguousarray(arr, NP_OF[t])
        o = self.host.vector(t, arr.shape[0])
        if arr.shape[0]:
            C.memmove(o + 16, arr.ctypes.data, arr.nbytes)
        return o

    def atom(self, t, v):
        o = self.host.atom(t)
        a = np.array([v], NP_OF[t])
        C.memmove(o + 8, a.ctypes.data, a.nbytes)
        return o

    def obj(self, t, x):
        """numpy array / list -> vector, scalar -> atom"""
        return self.atom(t, x) if np.ndim(x) == 0 else self.vec(t, x)

    def list_of(self, items):
        """a LIST object owning `items`"""
        o = self.host.vector(LIST, len(items))
        for i, it in enumerate(items):
            C.c_void_p.from_address(o + 16 + 8 * i).value = it
        return o

    def parted(self, t, parts):
        """a PARTED column (type 77 + t, core/rayforce.h:70-82): the per-partition vectors of a parted table"""
        o = self.list_of([self.vec(t, p) for p in parts])
        C.c_int8.from_address(o + 2).value = 77 + t
        return o

    def parted_index(self, groups, filt=None):
        """[INDEX_TYPE_PARTEDCOMMON, groups, -, -, -, filter, -] (core/math.c:1898-1899, core/index.c:1696-1699)"""
        items = [self.atom(capi.I64, 2), self.atom(capi.I64, groups), self.NULL, self.atom(capi.I64, capi.NULL_I64), self.NULL,
                 filt if filt is not None else self.NULL, self.NULL]
        return self.list_of(items)

    def table(self, names, cols):
        """a TABLE (type 98, core/rayforce.c:325-334): [SYMBOL vector of column names (symbol ids), LIST of column vectors]"""
        o = self.list_of([self.vec(capi.SYMBOL, np.asarray(names, np.int64)), self.list_of(cols)])
        C.c_int8.from_address(o + 2).value = 98
        return o

    @staticmethod
    def attrs_of(o):
        return C.c_uint8.from_address(o + 3).value

    def drop(self, *objs):
        for o in objs:
            if o:
                self.host.drop_obj(o)

    @staticmethod
    def type_of(o):
        return C.c_int8.from_address(o + 2).value

    @staticmethod
    def len_of(o):
        return C.c_int64.from_address(o + 8).value

    def items(self, o):
        return [C.c_void_p.from_address(o + 16 + 8 * i).value for i in range(self.len_of(o))]

    def value(self, o, drop=True):
        """-> (numpy scalar or array copy, type code); raises OpsError / Declined"""
        if not o:
            raise Declined()
        t = self.type_of(o)
        if t == T_ERR:
            raise OpsError(self.L.rfb_ops_builtin_last_error().decode())
        if t < 0:
            dt = np.dtype(NP_OF[-t])
            v = np.frombuffer((C.c_char * dt.itemsize).from_address(o + 8), dtype=dt)[0].copy()
            t = -t
        else:
            dt, n = np.dtype(NP_OF[t]), self.len_of(o)
            v = np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(o + 16), dtype=dt).copy() if n else np.empty(0, dt)
        if drop:
            self.drop(o)
        return v, t

    # ---- calls: op("ray_lt", x_obj, y_obj) -> raw object pointer
    def call(self, name, *args):
        return getattr(self.L, "rfb_" + name)(*args)

    def where_fold(self, cmp_op, fold, pred, k, val):
        args = (C.c_void_p * 5)(self.atom(capi.I64, cmp_op), self.atom(capi.I64, fold), pred, k, val)
        r = self.L.rfb_where_fold(args, 5)
        self.drop(args[0], args[1])
        return r

    def scope(self):
        ops = self

        class _S:
            def __enter__(self_inner):
                ops.L.rfb_ops_scope_begin()

            def __exit__(self_inner, *a):
                ops.L.rfb_ops_scope_end()
        return _S()

    @property
    def launches(self):
        return int(self.L.rfb_ops_launches())
