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
    """north_star tolerance for fp64 reductions: within 1 ULP.  The reference CPU path's own summation order is unspecified
    (thread count x -fassociative-math), so "of the reference" is not well defined; the GPU kernels accumulate error-free
    (TwoSum) and round once, and are held to 1 ULP of the EXACT sum (a __float128 sum rounded to double) — which also
    puts them within the reference's own error of the reference.  `cpu` is kept for reporting only."""
    return abs(got - exact) <= ulp(exact)


def same_f64(a, b, zero_sign=True, max_ulp=0) -> bool:
    """bit-equality with all NaNs identified (the reference treats any NaN as the null).
    zero_sign=False identifies -0.0 with +0.0: the reference is built with -funsafe-math-optimizations (which implies
    -fno-signed-zeros) and its own goldens print both as "0.0" (tests/lang.c:2551,2565 of the reference).
    max_ulp=1 is used for float division only: -freciprocal-math lets the reference's compiler turn x / atom into
    x * (1 / atom), so its quotients are only defined to 1 ULP (golden tests/lang.c:2395: (div [-3.0] -5.0))."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    a, b = a[~na], b[~nb]
    if not zero_sign:
        a, b = a + 0.0, b + 0.0          # -0.0 + 0.0 == +0.0
    if max_ulp == 0:
        return bool(np.array_equal(a.view(np.int64), b.view(np.int64)))
    fin = np.isfinite(a) & np.isfinite(b)
    if not np.array_equal(a[~fin], b[~fin]):
        return False
    return bool(np.all(np.abs(a[fin] - b[fin]) <= max_ulp * np.spacing(np.abs(b[fin]))))


ALL_ARITH_TYPES = [ob.B8, ob.U8, ob.I16, ob.I32, ob.I64, ob.DATE, ob.TIME, ob.TIMESTAMP, ob.F64]


def typed_col(t, n, seed):
    """small magnitudes (so products stay meaningful), zeros (division), nulls, and a few extreme values per type"""
    if t == ob.B8:
        return (np.random.default_rng(seed).integers(0, 2, n)).astype(np.uint8)
    if t == ob.U8:
        a = rng_col(t, n, seed, lo=0, hi=256)
        a[::17] = 0
        return a
    if t == ob.TIMESTAMP:
        a = rng_col(t, n, seed, null_frac=0.04, lo=-3 * 86_400_000_000_000, hi=40 * 86_400_000_000_000)
    elif t == ob.TIME:
        a = rng_col(t, n, seed, null_frac=0.04, lo=-1000, hi=86_400_000)
    else:
        a = rng_col(t, n, seed, null_frac=0.04, lo=-50, hi=50)
        if t == ob.F64:
            a = np.round(a * 4) / 4
    a[::17] = 0
    if t != ob.F64:
        info = np.iinfo(a.dtype)
        a[5::1013] = info.max
        a[7::1013] = info.min + 1
    return a
