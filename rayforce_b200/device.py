"""Host-side helper over the C ABI: one ``Context`` per GPU.  Device columns are torch CUDA tensors used purely as
HBM allocations (``data_ptr()`` is handed to the C ABI); host columns are numpy arrays.  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import capi
from .capi import RfbError, Fold, Scalar, check

NP_OF = {capi.B8: np.uint8, capi.U8: np.uint8, capi.I16: np.int16, capi.I32: np.int32, capi.DATE: np.int32,
         capi.TIME: np.int32, capi.I64: np.int64, capi.SYMBOL: np.int64, capi.TIMESTAMP: np.int64, capi.F64: np.float64}


def _dptr(t) -> int:
    """device pointer of a torch CUDA tensor / raw int / None"""
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    if not t.is_cuda:
        raise RfbError(capi.ERR_ARG, "expected a CUDA tensor (device layer takes device pointers)")
    if not t.is_contiguous():
        raise RfbError(capi.ERR_ARG, "device columns must be contiguous")
    return t.data_ptr()


def _hptr(a: np.ndarray) -> int:
    if not a.flags["C_CONTIGUOUS"]:
        raise RfbError(capi.ERR_ARG, "host columns must be contiguous")
    return a.ctypes.data


class ColumnFile:
    """A reference column file mapped read-only; `.array` is a numpy view of the payload (no copy)."""

    def __init__(self, path: str):
        self.lib = capi.load()
        self.f = capi.ColumnFile()
        check(self.lib.rfb_column_file_open(os.fsencode(path), C.byref(self.f)))
        self.type, self.len, self.attrs = self.f.type, self.f.len, self.f.attrs
        dt = np.dtype(NP_OF[self.type])
        buf = (C.c_char * (self.len * dt.itemsize)).from_address(self.f.payload) if self.len else b""
        self.array = np.frombuffer(buf, dtype=dt)

    def close(self):
        if self.f.map_base:
            self.array = None
            self.lib.rfb_column_file_close(C.byref(self.f))

    @staticmethod
    def write(path: str, type_: int, arr: np.ndarray, attrs: int = 0):
        arr = np.ascontiguousarray(arr, NP_OF[type_])
        check(capi.load().rfb_column_file_write(os.fsencode(path), type_, attrs, arr.ctypes.data, arr.shape[0]))


class FoldResult:
    """Python view of rfb_fold_t for a column of element type `type`."""

    def __init__(self, f: Fold, type_: int):
        self.type = type_
        self.rows, self.nonnull = f.rows, f.nonnull
        flt = type_ == capi.F64
        self.sum = f.sum_f64 if flt else f.sum_i64
        self.min = f.min_f64 if flt else f.min_i64
        self.max = f.max_f64 if flt else f.max_i64

    @property
    def avg(self) -> float:
        """ray_avg (reference core/math.c:2445-2526): sum / non-null count, 0Nf when nothing was counted."""
        if self.nonnull == 0:
            return float("nan")
        return float(self.sum) / float(self.nonnull)

    def __repr__(self):
        return "FoldResult(rows=%d nonnull=%d sum=%r min=%r max=%r)" % (self.rows, self.nonnull, self.sum, self.min, self.max)


class Context:
    """rfb_ctx_t wrapper.  `stream`: adopt a caller-owned CUDA stream (e.g. torch's current stream) so that
    torch.cuda.Event timing and torch allocations order correctly with the kernels."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = capi.load()
        if self.lib.rfb_device_count() <= 0:
            raise RfbError(capi.ERR_CUDA, "no CUDA device: rayforce_b200 has no CPU fallback")
        h = C.c_void_p()
        check(self.lib.rfb_ctx_create(device, C.byref(h)))
        self.h = h
        self.device = device
        if stream is not None:
            check(self.lib.rfb_ctx_set_stream(self.h, C.c_void_p(stream)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rfb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def sync(self):
        check(self.lib.rfb_sync(self.h))

    def set_result_ptr(self, t):
        """redirect fold results to device-visible memory (torch CUDA tensor of >= 8 int64) or None to undo"""
        check(self.lib.rfb_ctx_set_result_ptr(self.h, C.c_void_p(_dptr(t)) if t is not None else None))

    @property
    def launches(self) -> int:
        return int(self.lib.rfb_launch_count(self.h))

    @property
    def sm_count(self) -> int:
        return int(self.lib.rfb_ctx_sm_count(self.h))

    def fill_splitmix(self, type_: int, x, n: int, seed: int, modulus: int = 0, offset: int = 0, null_every: int = 0,
                      f64_scale: float = 1.0):
        check(self.lib.rfb_fill_splitmix_dev(self.h, type_, _dptr(x), n, seed, modulus, offset, null_every, f64_scale))

    # ---- device layer
    def fold(self, folds: int, type_: int, x, n: int) -> FoldResult:
        f = Fold()
        check(self.lib.rfb_fold_dev(self.h, folds, type_, _dptr(x), n, C.byref(f)))
        return FoldResult(f, type_)

    def filter_fold(self, cmp_op: int, pred_type: int, pred, k, folds: int, val_type: int, val, n: int,
                    k_type: int | None = None) -> FoldResult:
        s = Scalar.of(pred_type if k_type is None else k_type, k)
        f = Fold()
        check(self.lib.rfb_filter_fold_dev(self.h, cmp_op, pred_type, _dptr(pred), C.byref(s), folds, val_type,
                                           _dptr(val), n, C.byref(f)))
        return FoldResult(f, val_type)

    def filter_fold_async(self, cmp_op: int, pred_type: int, pred, k, folds: int, val_type: int, val, n: int):
        """enqueue only (no host synchronisation); collect with fold_result()"""
        s = Scalar.of(pred_type, k)
        check(self.lib.rfb_filter_fold_dev(self.h, cmp_op, pred_type, _dptr(pred), C.byref(s), folds, val_type,
                                           _dptr(val), n, None))

    def fold_result(self, type_: int) -> FoldResult:
        f = Fold()
        check(self.lib.rfb_fold_result(self.h, C.byref(f)))
        return FoldResult(f, type_)

    def peer_mailbox_setup(self, rank: int, world: int, group=None) -> None:
        """one process per GPU: create this rank's mailbox, exchange the CUDA IPC handles over torch.distributed, map the peers'"""
        import torch.distributed as dist
        h = (C.c_char * 64)()
        check(self.lib.rfb_peer_mailbox_create(self.h, h))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        blob = b"".join(handles)
        check(self.lib.rfb_peer_mailbox_bind(self.h, rank, world, C.c_char_p(blob)))
        dist.barrier(group=group)                      # every mailbox is mapped everywhere before the first exchange

    def fold_allreduce_peers(self, type_: int) -> FoldResult:
        """after an *_async fold launch: the merged result of all ranks (one tiny kernel over NVLink peer memory)"""
        f = Fold()
        check(self.lib.rfb_fold_allreduce_peers(self.h, type_, C.byref(f)))
        return FoldResult(f, type_)

    def fold_allreduce_peers_async(self, type_: int) -> None:
        """enqueue the exchange only; collect the last one with fold_peers_result()"""
        check(self.lib.rfb_fold_allreduce_peers(self.h, type_, None))

    def fold_peers_result(self, type_: int) -> FoldResult:
        f = Fold()
        check(self.lib.rfb_fold_peers_result(self.h, C.byref(f)))
        return FoldResult(f, type_)

    def peer_groups_setup(self, rank: int, world: int, capacity: int, group=None) -> None:
        """one process per GPU: this rank's group exchange buffer (capacity rows), IPC handles exchanged over torch.distributed"""
        import torch.distributed as dist
        h = (C.c_char * 64)()
        check(self.lib.rfb_peer_groups_create(self.h, capacity, h))
        handles = [None] * world
        dist.all_gather_object(handles, bytes(h.raw), group=group)
        check(self.lib.rfb_peer_groups_bind(self.h, rank, world, C.c_char_p(b"".join(handles))))
        dist.barrier(group=group)

    def group_merge_peers(self, keys, sums, counts, max_groups: int):
        """the merged (keys, sums, counts) of all ranks' group lists in global first-occurrence order, over NVLink peer memory;
        raises RfbError(kind 'type') for a key domain that is not dense (the caller then gathers and re-groups)"""
        ok, os_, oc = self._empty(max_groups, capi.I64), self._empty(max_groups, capi.I64), self._empty(max_groups, capi.I64)
        g = C.c_int64(0)
        check(self.lib.rfb_group_merge_peers(self.h, _dptr(keys), _dptr(sums), _dptr(counts), keys.shape[0], _dptr(ok), _dptr(os_), _dptr(oc),
                                             max_groups, C.byref(g)))
        return ok[:g.value], os_[:g.value], oc[:g.value]

    def multi_filter_fold(self, preds, conjunction: bool, folds: int, val_type: int, val, n: int) -> FoldResult:
        """preds: [(cmp_op, type, column tensor, constant), ...] combined with and (True) / or (False)"""
        arr = (capi.Pred * len(preds))()
        for i, (op, t, col, k) in enumerate(preds):
            arr[i].op, arr[i].type, arr[i].col, arr[i].k = op, t, _dptr(col), Scalar.of(t, k)
        f = Fold()
        check(self.lib.rfb_multi_filter_fold_dev(self.h, len(preds), arr, int(conjunction), folds, val_type, _dptr(val), n, C.byref(f)))
        return FoldResult(f, val_type)

    def fma_fold_async(self, folds: int, a, b, c, n: int) -> None:
        """enqueue only; collect with fold_result() or fold_allreduce_peers()"""
        check(self.lib.rfb_fma_fold_dev(self.h, folds, _dptr(a), _dptr(b), _dptr(c), n, None))

    def fma_fold(self, folds: int, a, b, c, n: int) -> FoldResult:
        f = Fold()
        check(self.lib.rfb_fma_fold_dev(self.h, folds, _dptr(a), _dptr(b), _dptr(c), n, C.byref(f)))
        return FoldResult(f, capi.F64)

    def gather_fold(self, folds: int, type_: int, col, ids, m: int) -> FoldResult:
        f = Fold()
        check(self.lib.rfb_gather_fold_dev(self.h, folds, type_, _dptr(col), _dptr(ids), m, C.byref(f)))
        return FoldResult(f, type_)

    # ---- host layer (HOST pointers in, host results out; copies inside)
    def filter_fold_host(self, cmp_op: int, pred_type: int, pred: np.ndarray, k, folds: int, val_type: int,
                         val: np.ndarray, chunk_rows: int = 0):
        s = Scalar.of(pred_type, k)
        f = Fold()
        nb = C.c_int64(0)
        check(self.lib.rfb_filter_fold_host(self.h, cmp_op, pred_type, _hptr(pred), C.byref(s), folds, val_type,
                                            _hptr(val), pred.shape[0], chunk_rows, C.byref(f), C.byref(nb)))
        return FoldResult(f, val_type), nb.value

    def group_sum_count_host(self, key_type, keys, val, max_groups, cmp_op=None, pred_type=None, pred=None, k=None):
        """fused group-by over HOST columns -> (keys, sums, counts) numpy arrays, bytes copied host -> device"""
        ok, osum, oc = (np.empty(max(max_groups, 1), np.int64) for _ in range(3))
        g, nb = C.c_int64(0), C.c_int64(0)
        s = Scalar.of(pred_type, k) if pred is not None else None
        check(self.lib.rfb_group_sum_count_host(self.h, key_type, _hptr(keys), _hptr(val), keys.shape[0], cmp_op or 0, pred_type or 0,
                                                _hptr(pred) if pred is not None else None, C.byref(s) if s is not None else None, max_groups,
                                                _hptr(ok), _hptr(osum), _hptr(oc), C.byref(g), C.byref(nb)))
        return ok[:g.value], osum[:g.value], oc[:g.value], nb.value

    def fma_fold_host(self, folds: int, a: np.ndarray, b: np.ndarray, c: np.ndarray):
        f = Fold()
        nb = C.c_int64(0)
        check(self.lib.rfb_fma_fold_host(self.h, folds, _hptr(a), _hptr(b), _hptr(c), a.shape[0], C.byref(f), C.byref(nb)))
        return FoldResult(f, capi.F64), nb.value

    def fold_host(self, folds: int, type_: int, x: np.ndarray, chunk_rows: int = 0):
        f = Fold()
        nb = C.c_int64(0)
        check(self.lib.rfb_fold_host(self.h, folds, type_, _hptr(x), x.shape[0], chunk_rows, C.byref(f), C.byref(nb)))
        return FoldResult(f, type_), nb.value


class MultiGpu:
    """rfb_mgpu_*: every visible GPU from one host process (row-range shards of HOST columns, host-side merge of the partials)"""

    def __init__(self, ndev: int = 0):
        self.lib = capi.load()
        h = C.c_void_p()
        check(self.lib.rfb_mgpu_create(ndev, C.byref(h)))
        self.h = h

    @property
    def devices(self) -> int:
        return int(self.lib.rfb_mgpu_devices(self.h))

    def close(self):
        if self.h:
            self.lib.rfb_mgpu_destroy(self.h)
            self.h = None

    def filter_fold_host(self, cmp_op, pred_type, pred, k, folds, val_type, val, chunk_rows=0):
        s = Scalar.of(pred_type, k) if pred is not None else None
        f = Fold()
        nb = C.c_int64(0)
        check(self.lib.rfb_mgpu_filter_fold_host(self.h, cmp_op or 0, pred_type or 0, _hptr(pred) if pred is not None else None, _sref(s), folds,
                                                 val_type, _hptr(val), val.shape[0], chunk_rows, C.byref(f), C.byref(nb)))
        return FoldResult(f, val_type), nb.value

    def group_sum_count_host(self, key_type, keys, val, max_groups, cmp_op=None, pred_type=None, pred=None, k=None):
        ok, osum, oc = (np.empty(max(max_groups, 1), np.int64) for _ in range(3))
        g, nb = C.c_int64(0), C.c_int64(0)
        s = Scalar.of(pred_type, k) if pred is not None else None
        check(self.lib.rfb_mgpu_group_sum_count_host(self.h, key_type, _hptr(keys), _hptr(val), keys.shape[0], cmp_op or 0, pred_type or 0,
                                                     _hptr(pred) if pred is not None else None, _sref(s), max_groups, _hptr(ok), _hptr(osum),
                                                     _hptr(oc), C.byref(g), C.byref(nb)))
        return ok[:g.value], osum[:g.value], oc[:g.value], nb.value


# ---------------------------------------------------------------------------------------------------------------------
# Operator-exact device-layer building blocks.  Operands are torch CUDA tensors (vectors) or Python scalars (atoms).

def _operand(x, t):
    """-> (device pointer, length or -1, Scalar or None)"""
    import torch
    if isinstance(x, torch.Tensor):
        return _dptr(x), x.shape[0], None
    return 0, -1, Scalar.of(t, x)


def _torch_dtype(t):
    import torch
    return {capi.B8: torch.uint8, capi.U8: torch.uint8, capi.I16: torch.int16, capi.I32: torch.int32, capi.DATE: torch.int32,
            capi.TIME: torch.int32, capi.I64: torch.int64, capi.SYMBOL: torch.int64, capi.TIMESTAMP: torch.int64,
            capi.F64: torch.float64}[t]


def _sref(s):
    return C.byref(s) if s is not None else None


class _Ops:
    """mixin for Context: ray_eq.., ray_where, filter_collect, ray_add.., index_group, aggr_*, ray_sort_* on device columns"""

    def _empty(self, n, t):
        import torch
        with torch.cuda.device(self.device):
            return torch.empty(max(n, 0), dtype=_torch_dtype(t), device="cuda:%d" % self.device)

    def cmp(self, op, xt, x, yt, y):
        """ray_eq/ne/lt/gt/le/ge: -> B8 mask (uint8 tensor)"""
        xp, xn, xs = _operand(x, xt)
        yp, yn, ys = _operand(y, yt)
        n = xn if xn >= 0 else yn
        mask = self._empty(n, capi.B8)
        check(self.lib.rfb_cmp_dev(self.h, op, xt, xp, xn, _sref(xs), yt, yp, yn, _sref(ys), _dptr(mask)))
        self.sync()   # the context's stream is non-blocking: make the result visible to torch's streams
        return mask

    def mask_logic(self, op, a, b=None):
        """and / or (b: mask tensor or a Python bool broadcast) / not on B8 masks"""
        import torch
        out = self._empty(a.shape[0], capi.B8)
        if isinstance(b, torch.Tensor):
            check(self.lib.rfb_mask_logic_dev(self.h, op, _dptr(a), a.shape[0], _dptr(b), b.shape[0], 0, _dptr(out)))
        else:
            check(self.lib.rfb_mask_logic_dev(self.h, op, _dptr(a), a.shape[0], None, -1, int(bool(b)), _dptr(out)))
        self.sync()
        return out

    def where(self, mask):
        """ray_where: B8 mask -> ascending i64 row ids"""
        n = mask.shape[0]
        ids = self._empty(n, capi.I64)
        cnt = C.c_int64(0)
        check(self.lib.rfb_where_dev(self.h, _dptr(mask), n, _dptr(ids), C.byref(cnt)))
        return ids[:cnt.value]

    def cmp_where(self, op, t, x, k):
        n = x.shape[0]
        ids = self._empty(n, capi.I64)
        cnt = C.c_int64(0)
        s = Scalar.of(t, k)
        check(self.lib.rfb_cmp_where_dev(self.h, op, t, _dptr(x), n, C.byref(s), _dptr(ids), C.byref(cnt)))
        return ids[:cnt.value]

    def gather(self, t, col, ids):
        """filter_collect / at_ids"""
        out = self._empty(ids.shape[0], t)
        check(self.lib.rfb_gather_dev(self.h, t, _dptr(col), _dptr(ids), ids.shape[0], _dptr(out)))
        self.sync()
        return out

    def binop_type(self, op, xt, yt):
        r = self.lib.rfb_binop_type(op, xt, yt)
        if r < 0:
            raise RfbError(r, "binop %d: unsupported operand types %d, %d" % (op, xt, yt))
        return r

    def binop(self, op, xt, x, yt, y):
        """ray_add/sub/mul/div/fdiv/mod -> (tensor, result type)"""
        xp, xn, xs = _operand(x, xt)
        yp, yn, ys = _operand(y, yt)
        ot = self.lib.rfb_binop_type_form(op, 0 if xn >= 0 and yn >= 0 else 1 if xn >= 0 else 2, xt, yt)
        if ot < 0:
            raise RfbError(ot, "binop %d: unsupported operand types %d, %d" % (op, xt, yt))
        if xn >= 0 and yn >= 0 and xn != yn:
            check(self.lib.rfb_binop_dev(self.h, op, xt, xp, xn, _sref(xs), yt, yp, yn, _sref(ys), None))
        out = self._empty(xn if xn >= 0 else yn, ot)
        check(self.lib.rfb_binop_dev(self.h, op, xt, xp, xn, _sref(xs), yt, yp, yn, _sref(ys), _dptr(out)))
        self.sync()
        return out, ot

    def unop_f64(self, op, x):
        out = self._empty(x.shape[0], capi.F64)
        check(self.lib.rfb_unop_f64_dev(self.h, op, _dptr(x), x.shape[0], _dptr(out)))
        self.sync()
        return out

    def group_i64(self, keys, filt=None):
        """index_group on an I64 key column -> (group_ids, first_ids, GroupInfo)"""
        n = keys.shape[0] if filt is None else filt.shape[0]
        gids = self._empty(n, capi.I64)
        firsts = self._empty(n, capi.I64)
        info = capi.GroupInfo()
        check(self.lib.rfb_group_i64_dev(self.h, _dptr(keys), _dptr(filt), n, _dptr(gids), _dptr(firsts), C.byref(info)))
        return gids, firsts[:info.groups], info

    def group_keys(self, cols, filt=None):
        """index_group_list (perfect-hash fusion): group by the tuple of I64-kind key columns"""
        n = cols[0].shape[0] if filt is None else filt.shape[0]
        gids = self._empty(n, capi.I64)
        firsts = self._empty(n, capi.I64)
        info = capi.GroupInfo()
        arr = (C.c_void_p * len(cols))(*[_dptr(c) for c in cols])
        check(self.lib.rfb_group_keys_i64_dev(self.h, len(cols), arr, _dptr(filt), n, _dptr(gids), _dptr(firsts), C.byref(info)))
        return gids, firsts[:info.groups], info

    def aggr(self, op, vt, val, gids, groups, filt=None):
        """aggr_sum/min/max/count/avg -> (tensor[groups], result type)"""
        self.lib.rfb_options_reload()      # tests / sweeps switch strategies through the environment between calls
        ot = self.lib.rfb_aggr_type(op, vt)
        if ot < 0:
            raise RfbError(ot, "aggr %d: unsupported value type %d" % (op, vt))
        out = self._empty(groups, ot)
        check(self.lib.rfb_aggr_dev(self.h, op, vt, _dptr(val), _dptr(filt), _dptr(gids), gids.shape[0], groups, _dptr(out)))
        self.sync()
        return out, ot

    def aggr_last(self, vt, val, gids, groups, nchunks=1, filt=None):
        """aggr_last as the reference computes it on `nchunks` worker chunks -> (tensor[groups], result type)"""
        ot = self.lib.rfb_aggr_type(capi.A_LAST, vt)
        if ot < 0:
            raise RfbError(ot, "aggr_last: unsupported value type %d" % vt)
        out = self._empty(groups, ot)
        check(self.lib.rfb_aggr_last_dev(self.h, vt, _dptr(val), _dptr(filt), _dptr(gids), gids.shape[0], groups, nchunks, _dptr(out)))
        self.sync()
        return out, ot

    def group_rows(self, gids, groups, filt=None):
        """aggr_row / aggr_collect layout -> (row ids grouped by gid in row order, offsets[groups+1])"""
        n = gids.shape[0]
        rows, offs = self._empty(n, capi.I64), self._empty(groups + 1, capi.I64)
        check(self.lib.rfb_group_rows_dev(self.h, _dptr(gids), _dptr(filt), n, groups, _dptr(rows), _dptr(offs)))
        self.sync()
        return rows, offs

    def med(self, t, x) -> float:
        """ray_med (ungrouped)"""
        out = C.c_double(0.0)
        check(self.lib.rfb_med_dev(self.h, t, _dptr(x), x.shape[0], C.byref(out)))
        return out.value

    def stddev(self, t, x) -> float:
        """ray_dev (ungrouped)"""
        out = C.c_double(0.0)
        check(self.lib.rfb_stddev_dev(self.h, t, _dptr(x), x.shape[0], C.byref(out)))
        return out.value

    def find_rows(self, build_cols, probe_cols):
        """ray_find / index_left_join_obj: for every probe row the first build row with an equal key tuple, else NULL_I64"""
        nb, np_ = build_cols[0].shape[0], probe_cols[0].shape[0]
        ids = self._empty(np_, capi.I64)
        b = (C.c_void_p * len(build_cols))(*[_dptr(c) for c in build_cols])
        p = (C.c_void_p * len(probe_cols))(*[_dptr(c) for c in probe_cols])
        check(self.lib.rfb_find_rows_dev(self.h, len(build_cols), b, nb, p, np_, _dptr(ids)))
        self.sync()
        return ids

    def inner_join(self, build_cols, probe_cols):
        """index_inner_join_obj -> (probe row ids, build row ids) of the matching pairs"""
        nb, np_ = build_cols[0].shape[0], probe_cols[0].shape[0]
        pi, bi = self._empty(np_, capi.I64), self._empty(np_, capi.I64)
        b = (C.c_void_p * len(build_cols))(*[_dptr(c) for c in build_cols])
        p = (C.c_void_p * len(probe_cols))(*[_dptr(c) for c in probe_cols])
        cnt = C.c_int64(0)
        check(self.lib.rfb_inner_join_dev(self.h, len(build_cols), b, nb, p, np_, _dptr(pi), _dptr(bi), C.byref(cnt)))
        self.sync()
        return pi[:cnt.value], bi[:cnt.value]

    def asof_join(self, build_cols, time_type, build_time, probe_cols, probe_time):
        """index_asof_join_obj: last build row of the probe row's key with time <= the probe time, else NULL_I64"""
        nb, np_ = build_cols[0].shape[0], probe_cols[0].shape[0]
        ids = self._empty(np_, capi.I64)
        b = (C.c_void_p * len(build_cols))(*[_dptr(c) for c in build_cols])
        p = (C.c_void_p * len(probe_cols))(*[_dptr(c) for c in probe_cols])
        check(self.lib.rfb_asof_join_dev(self.h, len(build_cols), b, time_type, _dptr(build_time), nb, p, _dptr(probe_time), np_, _dptr(ids)))
        self.sync()
        return ids

    def distinct(self, keys):
        """ray_distinct on a dense I64-kind key column -> the distinct keys, ascending (RfbError kind "arg" when not dense)"""
        n = keys.shape[0]
        out = self._empty(n, capi.I64)
        cnt = C.c_int64(0)
        check(self.lib.rfb_distinct_i64_dev(self.h, _dptr(keys), n, _dptr(out), C.byref(cnt)))
        self.sync()
        return out[:cnt.value]

    def window_join(self, op, vt, val, right_cols, right_time, left_cols, win_lo, win_hi, jtype):
        """window-join / window-join1 aggregate over a right table ordered by (key, time) -> (tensor[len(left)], type)"""
        nr, nl = right_cols[0].shape[0], left_cols[0].shape[0]
        ot = capi.I64 if op == capi.A_COUNT else (capi.F64 if op == capi.A_AVG else vt)
        out = self._empty(nl, ot)
        r = (C.c_void_p * len(right_cols))(*[_dptr(c) for c in right_cols])
        l = (C.c_void_p * len(left_cols))(*[_dptr(c) for c in left_cols])
        check(self.lib.rfb_window_join_dev(self.h, len(right_cols), r, _dptr(right_time), nr, l, nl, _dptr(win_lo), _dptr(win_hi), jtype, op, vt,
                                           _dptr(val), _dptr(out)))
        self.sync()
        return out, ot

    def group_sum_count(self, key_type, keys, val, max_groups, cmp_op=None, pred_type=None, pred=None, k=None):
        """fused select {s: (sum v) c: (count v) from t by k [where (cmp p k)]} -> (keys, sums, counts) tensors"""
        self.lib.rfb_options_reload()      # tests / sweeps switch strategies through the environment between calls
        n = keys.shape[0]
        ok, osum, oc = (self._empty(max_groups, capi.I64) for _ in range(3))
        g = C.c_int64(0)
        s = Scalar.of(pred_type, k) if pred is not None else None
        check(self.lib.rfb_group_sum_count_dev(self.h, key_type, _dptr(keys), _dptr(val), n, cmp_op or 0, pred_type or 0,
                                               _dptr(pred), _sref(s), max_groups, _dptr(ok), _dptr(osum), _dptr(oc), C.byref(g)))
        return ok[:g.value], osum[:g.value], oc[:g.value]

    def sort(self, t, x, descending=False):
        """ray_sort_asc / ray_sort_desc -> i64 permutation"""
        perm = self._empty(x.shape[0], capi.I64)
        self.lib.rfb_options_reload()      # tests / sweeps switch the pass structure (RFB_SORT_ALGO) between calls
        check(self.lib.rfb_sort_dev(self.h, t, _dptr(x), x.shape[0], int(descending), _dptr(perm)))
        self.sync()
        return perm


for _name, _fn in list(vars(_Ops).items()):
    if not _name.startswith("__"):
        setattr(Context, _name, _fn)
