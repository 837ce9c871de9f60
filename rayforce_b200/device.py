"""Host-side helper over the C ABI: one ``Context`` per GPU.  Device columns are torch CUDA tensors used purely as
HBM allocations (``data_ptr()`` is handed to the C ABI); host columns are numpy arrays.  No arithmetic happens here.
"""
from __future__ import annotations

import ctypes as C

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

    def fold_host(self, folds: int, type_: int, x: np.ndarray, chunk_rows: int = 0):
        f = Fold()
        nb = C.c_int64(0)
        check(self.lib.rfb_fold_host(self.h, folds, type_, _hptr(x), x.shape[0], chunk_rows, C.byref(f), C.byref(nb)))
        return FoldResult(f, type_), nb.value
