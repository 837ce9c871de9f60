"""shared helpers for the parity tests (test infrastructure)"""
from __future__ import annotations

import math

import numpy as np

from oracle import bindings as ob

SIZES = [0, 1, 2, 15, 16, 17, 255, 4097, 16383, 16384, 16385, 100_003, 1_000_003]   # straddles POOL_SPLIT_THRESHOLD (16384)


def dev(a):
    """numpy array -> torch CUDA tensor (an HBM allocation; no torch arithmetic is used in the tests)"""
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


def rng_col(t, n, seed, null_frac=0.0, lo=None, hi=None):
    """seeded column of reference type t with in-band nulls"""
    r = np.random.default_rng(seed)
    dt = ob.NP_OF[t]
    if t == ob.F64:
        a = r.uniform(-1e3 if lo is None else lo, 1e3 if hi is None else hi, n)
        if null_frac:
            a[r.random(n) < null_frac] = np.nan
        return a
    info = np.iinfo(dt)
    lo = max(info.min + 1, -(1 << 40)) if lo is None else lo
    hi = min(info.max, 1 << 40) if hi is None else hi
    a = r.integers(lo, hi, n, dtype=np.int64).astype(dt)
    if null_frac and t not in (ob.U8, ob.B8):
        a[r.random(n) < null_frac] = info.min
    return a


def ulp(x: float) -> float:
    return math.ulp(x) if math.isfinite(x) else float("inf")


def f64_sum_ok(got: float, cpu: float, exact: float) -> bool:
    """north_star tolerance for fp64 reductions: within 1 ULP of the exactly-rounded sum, or no worse than the
    reference CPU path's own error (whose order is unspecified under -fassociative-math)."""
    return abs(got - exact) <= max(ulp(exact), abs(cpu - exact))


def same_f64(a, b) -> bool:
    """bit-equality with all NaNs identified (the reference treats any NaN as the null)"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    return bool(np.array_equal(na, nb) and np.array_equal(a[~na].view(np.int64), b[~nb].view(np.int64)))
