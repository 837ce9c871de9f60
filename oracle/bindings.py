"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

* ``Oracle``    -> oracle/librf_oracle.so   (rf_oracle.c, the plain-C restatement)
* ``Reference`` -> oracle/_ref/librayforce_ref.so (the unmodified reference compiled from source by oracle/Makefile)

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.  The product package
(rayforce_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "librf_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "librayforce_ref.so")
REF_BIN = os.path.join(HERE, "_ref", "rayforce_ref")

# reference type codes (core/rayforce.h:50-62)
B8, U8, I16, I32, I64, SYMBOL, DATE, TIME, TIMESTAMP, F64 = 1, 2, 3, 4, 5, 6, 7, 8, 9, 10
EQ, NE, LT, GT, LE, GE = range(6)
SUM, MIN, MAX, CNT, AVG, COUNT, MED, DEV, FIRST, LAST = range(10)
ADD, SUB, MUL, DIV, FDIV, MOD, XBAR = range(7)
ROUND, FLOOR, CEIL = range(3)
ATOM = -1
NULL_I16 = -(2 ** 15)
NULL_I32 = -(2 ** 31)
NULL_I64 = -(2 ** 63)
INF_I64 = 2 ** 63 - 1

NP_OF = {B8: np.uint8, U8: np.uint8, I16: np.int16, I32: np.int32, DATE: np.int32, TIME: np.int32,
         I64: np.int64, SYMBOL: np.int64, TIMESTAMP: np.int64, F64: np.float64}


def build(ref: bool = True) -> None:
    """Compile the restatement (always) and the reference (when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir("/root/reference/core"):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _operand(a, t):
    """-> (contiguous numpy array, length-or-ATOM).  Python scalars / 0-d arrays are atoms."""
    arr = np.asarray(a, dtype=NP_OF[t])
    if arr.ndim == 0:
        return arr.reshape(1).copy(), ATOM
    arr = np.ascontiguousarray(arr)
    return arr, arr.shape[0]


class OracleError(Exception):
    def __init__(self, code):
        super().__init__({-1: "type", -2: "length"}.get(code, str(code)))
        self.code = code


class Oracle:
    """Thin numpy facade over rf_oracle.h."""

    class GroupInfo(C.Structure):
        _fields_ = [("index_type", C.c_int), ("dense", C.c_int), ("groups", C.c_int64), ("min", C.c_int64),
                    ("max", C.c_int64), ("range", C.c_int64)]

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        vp, i64, ci = C.c_void_p, C.c_int64, C.c_int
        L.rfo_cmp.restype = i64
        L.rfo_cmp.argtypes = [ci, ci, vp, i64, ci, vp, i64, vp]
        L.rfo_where.restype = i64
        L.rfo_where.argtypes = [vp, i64, vp]
        L.rfo_at_ids.restype = ci
        L.rfo_at_ids.argtypes = [ci, vp, vp, i64, vp]
        L.rfo_fold.restype = ci
        L.rfo_fold.argtypes = [ci, ci, vp, i64, vp, C.POINTER(ci)]
        L.rfo_sum_f64_exact.restype = C.c_double
        L.rfo_sum_f64_exact.argtypes = [vp, i64]
        L.rfo_binop_type.restype = ci
        L.rfo_binop_type.argtypes = [ci, ci, ci]
        L.rfo_binop_form.restype = ci
        L.rfo_binop_form.argtypes = [ci, ci, ci, ci]
        L.rfo_binop.restype = i64
        L.rfo_binop.argtypes = [ci, ci, vp, i64, ci, vp, i64, vp, C.POINTER(ci)]
        L.rfo_unop_f64.restype = ci
        L.rfo_unop_f64.argtypes = [ci, vp, i64, vp]
        L.rfo_group_i64.restype = ci
        L.rfo_group_i64.argtypes = [vp, vp, i64, vp, vp, vp, C.POINTER(Oracle.GroupInfo)]
        L.rfo_aggr.restype = ci
        L.rfo_aggr.argtypes = [ci, ci, vp, vp, vp, i64, i64, vp, C.POINTER(ci)]
        L.rfo_sort.restype = ci
        L.rfo_sort.argtypes = [ci, vp, i64, ci, vp]
        L.rfo_group_rows.restype = ci
        L.rfo_group_rows.argtypes = [vp, vp, i64, i64, vp, vp]
        for f in (L.rfo_med, L.rfo_dev):
            f.restype = ci
            f.argtypes = [ci, vp, i64, C.POINTER(C.c_double)]
        L.rfo_splitmix64.restype = C.c_uint64
        L.rfo_splitmix64.argtypes = [C.c_uint64, C.c_uint64]

    def cmp(self, op, xt, x, yt, y):
        xa, xn = _operand(x, xt)
        ya, yn = _operand(y, yt)
        n = xn if xn >= 0 else (yn if yn >= 0 else 1)
        out = np.empty(max(n, 1), np.uint8)
        r = self.L.rfo_cmp(op, xt, _ptr(xa), xn, yt, _ptr(ya), yn, _ptr(out))
        if r < 0:
            raise OracleError(r)
        return out[:r]

    def where(self, mask):
        mask = np.ascontiguousarray(mask, np.uint8)
        ids = np.empty(mask.shape[0], np.int64)
        c = self.L.rfo_where(_ptr(mask), mask.shape[0], _ptr(ids))
        return ids[:c].copy()

    def at_ids(self, t, col, ids):
        col = np.ascontiguousarray(col, NP_OF[t])
        ids = np.ascontiguousarray(ids, np.int64)
        out = np.empty(ids.shape[0], NP_OF[t])
        r = self.L.rfo_at_ids(t, _ptr(col), _ptr(ids), ids.shape[0], _ptr(out))
        if r < 0:
            raise OracleError(r)
        return out

    def fold(self, op, t, x):
        """-> (python/numpy scalar, result type code)"""
        x = np.ascontiguousarray(x, NP_OF[t])
        out = np.zeros(1, np.int64)
        ot = C.c_int(0)
        r = self.L.rfo_fold(op, t, _ptr(x), x.shape[0], _ptr(out), C.byref(ot))
        if r < 0:
            raise OracleError(r)
        return out.view(NP_OF[ot.value])[0], ot.value

    def sum_f64_exact(self, x):
        x = np.ascontiguousarray(x, np.float64)
        return self.L.rfo_sum_f64_exact(_ptr(x), x.shape[0])

    def binop_type(self, op, xt, yt):
        return self.L.rfo_binop_type(op, xt, yt)

    def binop_form(self, op, form, xt, yt):
        return self.L.rfo_binop_form(op, form, xt, yt)

    def binop(self, op, xt, x, yt, y):
        xa, xn = _operand(x, xt)
        ya, yn = _operand(y, yt)
        ot = self.L.rfo_binop_type(op, xt, yt)
        if ot < 0:      # outside I32/I64/F64: the full type matrix, per operand form
            ot = self.binop_form(op, 0 if xn >= 0 and yn >= 0 else 1 if xn >= 0 else 2, xt, yt)
        if ot < 0:
            raise OracleError(ot)
        n = xn if xn >= 0 else (yn if yn >= 0 else 1)
        out = np.empty(max(n, 1), NP_OF[ot])
        o2 = C.c_int(0)
        r = self.L.rfo_binop(op, xt, _ptr(xa), xn, yt, _ptr(ya), yn, _ptr(out), C.byref(o2))
        if r < 0:
            raise OracleError(r)
        return out[:r], ot

    def unop_f64(self, op, x):
        x = np.ascontiguousarray(x, np.float64)
        out = np.empty_like(x)
        self.L.rfo_unop_f64(op, _ptr(x), x.shape[0], _ptr(out))
        return out

    def group_i64(self, keys, filt=None):
        keys = np.ascontiguousarray(keys, np.int64)
        if filt is not None:
            filt = np.ascontiguousarray(filt, np.int64)
        n = keys.shape[0] if filt is None else filt.shape[0]
        gids = np.empty(n, np.int64)
        firsts = np.empty(n, np.int64)
        info = Oracle.GroupInfo()
        r = self.L.rfo_group_i64(_ptr(keys), _ptr(filt), n, _ptr(gids), _ptr(firsts), None, C.byref(info))
        if r < 0:
            raise OracleError(r)
        return gids, firsts[:info.groups].copy(), info

    def group_multi(self, cols, filt=None):
        cols = [np.ascontiguousarray(c, np.int64) for c in cols]
        if filt is not None:
            filt = np.ascontiguousarray(filt, np.int64)
        n = cols[0].shape[0] if filt is None else filt.shape[0]
        gids, firsts = np.empty(n, np.int64), np.empty(n, np.int64)
        arr = (C.c_void_p * len(cols))(*[c.ctypes.data for c in cols])
        g = C.c_int64(0)
        self.L.rfo_group_multi.restype = C.c_int
        self.L.rfo_group_multi.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        r = self.L.rfo_group_multi(len(cols), arr, _ptr(filt), n, _ptr(gids), _ptr(firsts), C.byref(g))
        if r < 0:
            raise OracleError(r)
        return gids, firsts[:g.value].copy(), g.value

    def aggr(self, op, vt, val, gids, groups, filt=None):
        val = np.ascontiguousarray(val, NP_OF[vt])
        gids = np.ascontiguousarray(gids, np.int64)
        if filt is not None:
            filt = np.ascontiguousarray(filt, np.int64)
        out = np.zeros(max(groups, 1), np.int64)
        ot = C.c_int(0)
        r = self.L.rfo_aggr(op, vt, _ptr(val), _ptr(filt), _ptr(gids), gids.shape[0], groups, _ptr(out), C.byref(ot))
        if r < 0:
            raise OracleError(r)
        dt = NP_OF[ot.value]
        return out.view(np.uint8)[: groups * np.dtype(dt).itemsize].view(dt).copy(), ot.value

    def aggr_last(self, vt, val, gids, groups, nchunks=1, filt=None):
        """aggr_last on `nchunks` worker chunks (core/aggr.c:851-1075) -> (array[groups], result type)"""
        val = np.ascontiguousarray(val, NP_OF[vt])
        gids = np.ascontiguousarray(gids, np.int64)
        if filt is not None:
            filt = np.ascontiguousarray(filt, np.int64)
        out = np.zeros(max(groups, 1), np.int64)
        ot = C.c_int(0)
        self.L.rfo_aggr_last.restype = C.c_int
        self.L.rfo_aggr_last.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]
        r = self.L.rfo_aggr_last(vt, _ptr(val), _ptr(filt), _ptr(gids), gids.shape[0], groups, nchunks, _ptr(out), C.byref(ot))
        if r < 0:
            raise OracleError(r)
        dt = NP_OF[ot.value]
        return out.view(np.uint8)[: groups * np.dtype(dt).itemsize].view(dt).copy(), ot.value

    def parted_aggr(self, op, vt, parts, combine):
        """PARTED_MAP without a filter -> (array of 1 or len(parts) values, result type)"""
        parts = [np.ascontiguousarray(p, NP_OF[vt]) for p in parts]
        pa = (C.c_void_p * len(parts))(*[p.ctypes.data for p in parts])
        lens = np.array([p.shape[0] for p in parts], np.int64)
        out = np.zeros(max(len(parts), 1), np.int64)
        ot = C.c_int(0)
        self.L.rfo_parted_aggr.restype = C.c_int
        self.L.rfo_parted_aggr.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        r = self.L.rfo_parted_aggr(op, vt, len(parts), pa, _ptr(lens), int(combine), _ptr(out), C.byref(ot))
        if r < 0:
            raise OracleError(r)
        dt = np.dtype(NP_OF[ot.value])
        cnt = 1 if combine else len(parts)
        return out.view(np.uint8)[: cnt * dt.itemsize].view(dt).copy(), ot.value

    def group_rows(self, gids, groups, filt=None):
        """aggr_row / aggr_collect layout: (rows grouped by gid in push order, offsets[groups+1])"""
        gids = np.ascontiguousarray(gids, np.int64)
        if filt is not None:
            filt = np.ascontiguousarray(filt, np.int64)
        rows, offs = np.empty(gids.shape[0], np.int64), np.empty(groups + 1, np.int64)
        self.L.rfo_group_rows(_ptr(gids), _ptr(filt), gids.shape[0], groups, _ptr(rows), _ptr(offs))
        return rows, offs

    def find_rows(self, build_cols, probe_cols):
        b = [np.ascontiguousarray(c, np.int64) for c in build_cols]
        p = [np.ascontiguousarray(c, np.int64) for c in probe_cols]
        ids = np.empty(p[0].shape[0], np.int64)
        ba = (C.c_void_p * len(b))(*[c.ctypes.data for c in b])
        pa = (C.c_void_p * len(p))(*[c.ctypes.data for c in p])
        self.L.rfo_find_rows.restype = C.c_int
        self.L.rfo_find_rows.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int64, C.POINTER(C.c_void_p), C.c_int64, C.c_void_p]
        self.L.rfo_find_rows(len(b), ba, b[0].shape[0], pa, p[0].shape[0], _ptr(ids))
        return ids

    def asof_join(self, build_cols, tt, build_time, probe_cols, probe_time):
        b = [np.ascontiguousarray(c, np.int64) for c in build_cols]
        p = [np.ascontiguousarray(c, np.int64) for c in probe_cols]
        bt, pt = np.ascontiguousarray(build_time, NP_OF[tt]), np.ascontiguousarray(probe_time, NP_OF[tt])
        ids = np.empty(p[0].shape[0], np.int64)
        ba = (C.c_void_p * len(b))(*[c.ctypes.data for c in b])
        pa = (C.c_void_p * len(p))(*[c.ctypes.data for c in p])
        self.L.rfo_asof_join.restype = C.c_int
        self.L.rfo_asof_join.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p, C.c_int64, C.c_void_p]
        r = self.L.rfo_asof_join(len(b), ba, tt, _ptr(bt), b[0].shape[0], pa, _ptr(pt), p[0].shape[0], _ptr(ids))
        if r < 0:
            raise OracleError(r)
        return ids

    def window_aggr(self, op, vt, val, right_cols, rtime, left_cols, wlo, whi, jtype):
        """window join aggregate over a right table ordered by (key, time) -> (array[len(left)], type)"""
        r = [np.ascontiguousarray(c, np.int64) for c in right_cols]
        l = [np.ascontiguousarray(c, np.int64) for c in left_cols]
        ra = (C.c_void_p * len(r))(*[c.ctypes.data for c in r])
        la = (C.c_void_p * len(l))(*[c.ctypes.data for c in l])
        ll, rl = l[0].shape[0], r[0].shape[0]
        first, last = np.empty(max(ll, 1), np.int64), np.empty(max(ll, 1), np.int64)
        L = self.L
        L.rfo_window_bounds.restype = C.c_int
        L.rfo_window_bounds.argtypes = [C.c_int, C.POINTER(C.c_void_p), C.c_int64, C.POINTER(C.c_void_p), C.c_int64, C.c_void_p, C.c_void_p]
        L.rfo_window_bounds(len(r), ra, rl, la, ll, _ptr(first), _ptr(last))
        val = np.ascontiguousarray(val, NP_OF[vt])
        rtime, wlo, whi = (np.ascontiguousarray(a, np.int32) for a in (rtime, wlo, whi))
        out = np.zeros(max(ll, 1), np.int64)
        ot = C.c_int(0)
        L.rfo_window_aggr.restype = C.c_int
        L.rfo_window_aggr.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_int)]
        rc = L.rfo_window_aggr(op, vt, _ptr(val), _ptr(rtime), ll, _ptr(first), _ptr(last), _ptr(wlo), _ptr(whi), jtype, _ptr(out), C.byref(ot))
        if rc < 0:
            raise OracleError(rc)
        return out[:ll].view(NP_OF[ot.value]).copy(), ot.value

    def inner_join(self, build_cols, probe_cols):
        ids = self.find_rows(build_cols, probe_cols)
        pi = np.nonzero(ids != NULL_I64)[0].astype(np.int64)
        return pi, ids[pi]

    def distinct(self, keys):
        keys = np.ascontiguousarray(keys, np.int64)
        out = np.empty(max(keys.shape[0], 1), np.int64)
        self.L.rfo_distinct_i64.restype = C.c_int64
        self.L.rfo_distinct_i64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        c = self.L.rfo_distinct_i64(_ptr(keys), keys.shape[0], _ptr(out))
        if c < 0:
            raise OracleError(-4)
        return out[:c].copy()

    def med(self, t, x):
        return self._stat(self.L.rfo_med, t, x)

    def dev(self, t, x):
        return self._stat(self.L.rfo_dev, t, x)

    def _stat(self, fn, t, x):
        x = np.ascontiguousarray(x, NP_OF[t])
        out = C.c_double(0.0)
        r = fn(t, _ptr(x), x.shape[0], C.byref(out))
        if r < 0:
            raise OracleError(r)
        return out.value

    def sort(self, t, x, descending=False):
        x = np.ascontiguousarray(x, NP_OF[t])
        perm = np.empty(x.shape[0], np.int64)
        r = self.L.rfo_sort(t, _ptr(x), x.shape[0], int(descending), _ptr(perm))
        if r < 0:
            raise OracleError(r)
        return perm

    def splitmix64(self, seed, i):
        return self.L.rfo_splitmix64(seed, i)


class RefError(Exception):
    pass


class Reference:
    """Operator-level access to the unmodified reference (SURVEY.md §8c "operator level from C").

    Objects are the reference's own obj_t (16-byte header + payload, core/rayforce.h:112-133); this class only
    builds vectors/atoms, calls the reference's exported operator functions and reads results back into numpy.
    One runtime per process (ray_init creates the thread pool with all cores unless RAYFORCE_CORES is... no such
    knob: the pool size is fixed at ray_init time, see `cores`).
    """
    _inst = None

    @classmethod
    def available(cls) -> bool:
        return os.path.exists(REF_SO)

    @classmethod
    def get(cls) -> "Reference":
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self, path: str = REF_SO):
        L = self.L = C.CDLL(path)
        vp, i64 = C.c_void_p, C.c_int64
        L.ray_init.restype = C.c_int
        if L.ray_init() != 0:
            raise RefError("ray_init failed")
        L.vector.restype = vp
        L.vector.argtypes = [C.c_int8, i64]
        L.i64.restype = vp
        L.i64.argtypes = [i64]
        L.i32.restype = vp
        L.i32.argtypes = [C.c_int32]
        L.f64.restype = vp
        L.f64.argtypes = [C.c_double]
        for name, ct in (("b8", C.c_uint8), ("u8", C.c_uint8), ("i16", C.c_int16), ("adate", C.c_int32), ("atime", C.c_int32),
                         ("timestamp", C.c_int64)):       # atom constructors, core/rayforce.c:150-235
            f = getattr(L, name)
            f.restype, f.argtypes = vp, [ct]
        L.drop_obj.restype = None
        L.drop_obj.argtypes = [vp]
        L.clone_obj.restype = vp
        L.clone_obj.argtypes = [vp]
        for name in ("ray_sum", "ray_min", "ray_max", "ray_cnt", "ray_avg", "ray_count", "ray_where", "ray_round",
                     "ray_floor", "ray_ceil", "ray_sort_asc", "ray_sort_desc", "ray_iasc", "ray_idesc", "ray_asc",
                     "ray_desc", "ray_med", "ray_dev", "ray_distinct"):
            f = getattr(L, name)
            f.restype = vp
            f.argtypes = [vp]
        for name in ("ray_eq", "ray_ne", "ray_lt", "ray_gt", "ray_le", "ray_ge", "ray_add", "ray_sub", "ray_mul",
                     "ray_div", "ray_fdiv", "ray_mod", "ray_xbar", "filter_map", "filter_collect", "index_group", "group_map",
                     "aggr_sum", "aggr_min", "aggr_max", "aggr_count", "aggr_avg", "aggr_first", "aggr_last", "aggr_med", "aggr_dev",
                     "aggr_row", "aggr_collect"):
            f = getattr(L, name)
            f.restype = vp
            f.argtypes = [vp, vp]
        L.eval_str.restype = vp
        L.eval_str.argtypes = [C.c_char_p]
        self.NULL_OBJ = C.addressof(C.c_char.in_dll(L, "__NULL_OBJ"))
        L.pool_get_executors_count.restype = i64
        L.pool_get_executors_count.argtypes = [vp]
        L.pool_get.restype = vp
        self.cores = int(L.pool_get_executors_count(L.pool_get()))

    # ---- object construction / inspection
    def vec(self, t, arr):
        arr = np.ascontiguousarray(arr, NP_OF[t])
        o = self.L.vector(t, arr.shape[0])
        if arr.shape[0]:
            C.memmove(o + 16, arr.ctypes.data, arr.nbytes)
        return o

    def make_list(self, items):
        """a reference LIST (type 0) owning the given objects"""
        o = self.L.vector(0, len(items))
        for i, it in enumerate(items):
            C.c_void_p.from_address(o + 16 + 8 * i).value = it
        return o

    def find(self, x, y):
        """ray_find(x, y): for every y the first index in x with that value, else null (I64 vectors)"""
        xo, yo = self.vec(I64, x), self.vec(I64, y)
        self.L.ray_find.restype, self.L.ray_find.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p]
        r = self.L.ray_find(xo, yo)
        if self.is_err(r):
            raise RefError("ray_find")
        out = self.to_numpy(r)[0]
        self.drop(xo, yo)
        return out

    def asof_index(self, lcols, tt, lx, rcols, rx):
        """index_asof_join_obj(lcols, lxcol, rcols, rxcol) on lists of I64 key columns + a time column on each side"""
        lo, ro = self.make_list([self.vec(I64, c) for c in lcols]), self.make_list([self.vec(I64, c) for c in rcols])
        lxo, rxo = self.vec(tt, lx), self.vec(tt, rx)
        f = self.L.index_asof_join_obj
        f.restype, f.argtypes = C.c_void_p, [C.c_void_p] * 4
        r = f(lo, lxo, ro, rxo)
        if self.is_err(r):
            raise RefError("index_asof_join_obj")
        out = self.to_numpy(r)[0]
        self.drop(lo, ro, lxo, rxo)
        return out

    def isin(self, x, y):
        """ray_in(x, y) on I64 vectors -> B8 mask"""
        xo, yo = self.vec(I64, x), self.vec(I64, y)
        self.L.ray_in.restype, self.L.ray_in.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p]
        r = self.L.ray_in(xo, yo)
        if self.is_err(r):
            raise RefError("ray_in")
        out = self.to_numpy(r)[0]
        self.drop(xo, yo)
        return out

    def join_index(self, lcols, rcols, inner=False):
        """index_left_join_obj / index_inner_join_obj on lists of I64 key columns (2+ columns: the hashed multi-column path)"""
        lo, ro = self.make_list([self.vec(I64, c) for c in lcols]), self.make_list([self.vec(I64, c) for c in rcols])
        f = self.L.index_inner_join_obj if inner else self.L.index_left_join_obj
        f.restype, f.argtypes = C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int64]
        r = f(lo, ro, len(lcols))
        if self.is_err(r):
            raise RefError("join index")
        if inner:
            its = self.list_items(r)
            out = (self.to_numpy(its[0], drop=False)[0].copy(), self.to_numpy(its[1], drop=False)[0].copy())
            self.drop(r)
        else:
            out = self.to_numpy(r)[0]
        self.drop(lo, ro)
        return out

    def vec_uninit(self, t, n):
        """reference-owned vector plus a numpy view over its payload (fill in place; no copy)"""
        o = self.L.vector(t, n)
        dt = np.dtype(NP_OF[t])
        buf = (C.c_char * (n * dt.itemsize)).from_address(o + 16)
        return o, np.frombuffer(buf, dtype=dt)

    def atom(self, t, v):
        if t == I64:
            return self.L.i64(int(v))
        if t == I32:
            return self.L.i32(int(v))
        if t == F64:
            return self.L.f64(float(v))
        ctor = {B8: "b8", U8: "u8", I16: "i16", DATE: "adate", TIME: "atime", TIMESTAMP: "timestamp"}.get(t)
        if ctor:
            return getattr(self.L, ctor)(int(v))
        raise RefError("atom type %d" % t)

    def operand(self, t, a):
        arr = np.asarray(a, dtype=NP_OF[t])
        return self.atom(t, arr) if arr.ndim == 0 else self.vec(t, arr)

    @staticmethod
    def type_of(o):
        return C.c_int8.from_address(o + 2).value

    @staticmethod
    def len_of(o):
        return C.c_int64.from_address(o + 8).value

    def is_err(self, o):
        return self.type_of(o) == 127

    def drop(self, *objs):
        for o in objs:
            if o:
                self.L.drop_obj(o)

    def to_numpy(self, o, drop=True):
        """atom -> (numpy scalar, type) ; vector -> (numpy array copy, type)"""
        t = self.type_of(o)
        if t == 127:
            raise RefError("reference returned an error object")
        if t < 0:
            dt = np.dtype(NP_OF[-t])
            v = np.frombuffer((C.c_char * dt.itemsize).from_address(o + 8), dtype=dt)[0].copy()
            if drop:
                self.drop(o)
            return v, -t
        n = self.len_of(o)
        dt = np.dtype(NP_OF[t])
        if n:
            out = np.frombuffer((C.c_char * (n * dt.itemsize)).from_address(o + 16), dtype=dt).copy()
        else:
            out = np.empty(0, dt)
        if drop:
            self.drop(o)
        return out, t

    def list_items(self, o):
        n = self.len_of(o)
        return [C.c_void_p.from_address(o + 16 + 8 * i).value for i in range(n)]

    def call1(self, name, x):
        return getattr(self.L, name)(x)

    def call2(self, name, x, y):
        return getattr(self.L, name)(x, y)

    # ---- numpy-level conveniences mirroring Oracle's API
    _CMP = ["ray_eq", "ray_ne", "ray_lt", "ray_gt", "ray_le", "ray_ge"]
    _BIN = ["ray_add", "ray_sub", "ray_mul", "ray_div", "ray_fdiv", "ray_mod", "ray_xbar"]
    _FOLD = ["ray_sum", "ray_min", "ray_max", "ray_cnt", "ray_avg", "ray_count", "ray_med", "ray_dev"]
    _AGGR = ["aggr_sum", "aggr_min", "aggr_max", None, "aggr_avg", "aggr_count", "aggr_med", "aggr_dev", "aggr_first", "aggr_last"]

    def _bin(self, name, xt, x, yt, y):
        xo, yo = self.operand(xt, x), self.operand(yt, y)
        r = self.call2(name, xo, yo)
        err = self.is_err(r)
        self.drop(xo, yo)
        if err:
            raise RefError(name)
        return self.to_numpy(r)

    def cmp(self, op, xt, x, yt, y):
        return self._bin(self._CMP[op], xt, x, yt, y)[0]

    def binop(self, op, xt, x, yt, y):
        return self._bin(self._BIN[op], xt, x, yt, y)

    def where(self, mask):
        m = self.vec(B8, mask)
        r = self.call1("ray_where", m)
        self.drop(m)
        return self.to_numpy(r)[0]

    def at_ids(self, t, col, ids):
        c, i = self.vec(t, col), self.vec(I64, ids)
        r = self.call2("filter_collect", c, i)
        self.drop(c, i)
        return self.to_numpy(r)[0]

    def fold(self, op, t, x):
        v = self.vec(t, x)
        r = self.call1(self._FOLD[op], v)
        err = self.is_err(r)
        self.drop(v)
        if err:
            raise RefError(self._FOLD[op])
        return self.to_numpy(r)

    def filter_fold(self, cmp_op, t, x, k, fold_op, vt=None, val=None):
        """the unfused operator pipeline of `select {(fold v) from t where (cmp x k)}`"""
        xo = self.vec(t, x)
        vo = xo if val is None else self.vec(vt, val)
        ko = self.atom(t, k)
        m = self.call2(self._CMP[cmp_op], xo, ko)
        ids = self.call1("ray_where", m)
        lazy = self.call2("filter_map", vo, ids)
        r = self.call1(self._FOLD[fold_op], lazy)
        out = self.to_numpy(r)
        self.drop(lazy, ids, m, ko, xo)
        if val is not None:
            self.drop(vo)
        return out

    def unop_f64(self, op, x):
        v = self.vec(F64, x)
        r = self.call1(["ray_round", "ray_floor", "ray_ceil"][op], v)
        self.drop(v)
        return self.to_numpy(r)[0]

    def sort(self, t, x, descending=False):
        v = self.vec(t, x)
        r = self.call1("ray_sort_desc" if descending else "ray_sort_asc", v)
        self.drop(v)
        return self.to_numpy(r)[0]

    def group_aggr(self, keys, vt, val, ops, filt=None):
        """index_group(keys, filter) then aggr_<op>(val, index) for each op.
        -> dict(groups, index_type, first_ids, results={op: (array, type)})"""
        ko = self.vec(I64, keys)
        vo = self.vec(vt, val)
        fo = self.NULL_OBJ if filt is None else self.vec(I64, filt)
        idx = self.call2("index_group", ko, fo)
        if self.is_err(idx):
            raise RefError("index_group")
        items = self.list_items(idx)
        out = {"index_type": C.c_int64.from_address(items[0] + 8).value,
               "groups": C.c_int64.from_address(items[1] + 8).value, "results": {}}
        meta = items[6]
        out["first_ids"] = self.to_numpy(meta, drop=False)[0] if self.type_of(meta) == I64 else None
        for op in ops:
            r = self.call2(self._AGGR[op], vo, idx)
            if self.is_err(r):
                raise RefError(self._AGGR[op])
            out["results"][op] = self.to_numpy(r)
        self.drop(idx, ko, vo)
        if filt is not None:
            self.drop(fo)
        return out

    def group_lists(self, keys, vt, val, filt=None):
        """index_group(keys, filter) then aggr_row / aggr_collect -> (row-id arrays, value arrays), one per group"""
        ko, vo = self.vec(I64, keys), self.vec(vt, val)
        fo = self.NULL_OBJ if filt is None else self.vec(I64, filt)
        idx = self.call2("index_group", ko, fo)
        if self.is_err(idx):
            raise RefError("index_group")
        out = []
        for name in ("aggr_row", "aggr_collect"):
            r = self.call2(name, vo, idx)
            if self.is_err(r):
                raise RefError(name)
            out.append([self.to_numpy(it, drop=False)[0].copy() for it in self.list_items(r)])
            self.drop(r)
        self.drop(idx, ko, vo)
        if filt is not None:
            self.drop(fo)
        return out[0], out[1]

    def eval(self, src: str):
        return self.L.eval_str(src.encode())
