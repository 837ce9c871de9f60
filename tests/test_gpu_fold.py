"""GPU parity: scan+fold kernels (k_fold.cu) vs the CPU oracle, through the C ABI.
Mirrors the reference's sum/min/max/avg/count goldens (tests/lang.c:2455-2540, 4065-4090) and its
`select ... where:` cases (tests/lang.c:2883-2897) at sizes straddling its 16384-row parallel threshold."""
import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200 import capi
from tests.util import SIZES, dev, rng_col, f64_sum_ok

pytestmark = pytest.mark.gpu

INT_TYPES = [ob.U8, ob.I16, ob.I32, ob.I64, ob.TIME]
CMPS = [ob.EQ, ob.NE, ob.LT, ob.GT, ob.LE, ob.GE]


def oracle_folds(O, t, col):
    """-> dict(sum, min, max, nonnull) from the oracle (None where the reference raises a type error)"""
    out = {}
    for name, op in (("sum", ob.SUM), ("min", ob.MIN), ("max", ob.MAX), ("nonnull", ob.CNT)):
        try:
            out[name] = O.fold(op, t, col)[0]
        except ob.OracleError:
            out[name] = None
    return out


def check_fold(got, want, t, n):
    assert got.rows == n
    if want["nonnull"] is not None:
        assert got.nonnull == int(want["nonnull"])
    if t == ob.F64:
        return
    if want["sum"] is not None:
        assert got.sum == int(want["sum"])
    if n > 0 or t != ob.U8:     # the oracle's U8 min/max of an empty vector is 0 by construction
        assert got.min == int(want["min"]) and got.max == int(want["max"])


@pytest.mark.parametrize("t", INT_TYPES)
@pytest.mark.parametrize("n", SIZES)
def test_fold_int(ctx, oracle, t, n):
    col = rng_col(t, n, seed=n + t, null_frac=0.05)
    got = ctx.fold(capi.F_ALL, t, dev(col) if n else None, n)
    check_fold(got, oracle_folds(oracle, t, col), t, n)


@pytest.mark.parametrize("t", [ob.I16, ob.I32, ob.I64])
def test_fold_all_null(ctx, oracle, t):
    col = np.full(5000, np.iinfo(ob.NP_OF[t]).min, ob.NP_OF[t])
    got = ctx.fold(capi.F_ALL, t, dev(col), col.shape[0])
    check_fold(got, oracle_folds(oracle, t, col), t, col.shape[0])
    assert got.nonnull == 0 and got.sum == 0 and got.min == np.iinfo(ob.NP_OF[t]).min


def test_fold_i32_sum_wraps_in_32_bits(ctx, oracle):
    # (sum [2147483647i 1i]) -> 0Ni in the reference
    col = np.array([2147483647, 1], np.int32)
    got = ctx.fold(capi.F_SUM, ob.I32, dev(col), 2)
    assert got.sum == int(oracle.fold(ob.SUM, ob.I32, col)[0]) == -(2 ** 31)


def test_fold_i64_wraps(ctx, oracle):
    col = np.array([2 ** 63 - 1, 5, 2 ** 63 - 1, 7] * 1000, np.int64)
    got = ctx.fold(capi.F_SUM, ob.I64, dev(col), col.shape[0])
    assert got.sum == int(oracle.fold(ob.SUM, ob.I64, col)[0])


@pytest.mark.parametrize("n", SIZES)
def test_fold_f64(ctx, oracle, n):
    col = rng_col(ob.F64, n, seed=n, null_frac=0.03)
    got = ctx.fold(capi.F_ALL, ob.F64, dev(col) if n else None, n)
    assert got.rows == n and got.nonnull == int(np.count_nonzero(~np.isnan(col)))
    cpu = float(oracle.fold(ob.SUM, ob.F64, col)[0])
    assert f64_sum_ok(got.sum, cpu, oracle.sum_f64_exact(col))
    wmin, wmax = float(oracle.fold(ob.MIN, ob.F64, col)[0]), float(oracle.fold(ob.MAX, ob.F64, col)[0])
    if got.nonnull:
        assert got.min == wmin and got.max == wmax
    else:
        assert np.isnan(got.min) and np.isnan(got.max) and np.isnan(wmin)


def test_fold_f64_exact_integers_bit_exact(ctx, oracle):
    # integer-valued doubles whose partial sums are all exactly representable: any order gives the same bits
    col = np.random.default_rng(7).integers(0, 1 << 20, 2_000_003).astype(np.float64)
    got = ctx.fold(capi.F_SUM, ob.F64, dev(col), col.shape[0])
    assert got.sum == float(oracle.fold(ob.SUM, ob.F64, col)[0]) == float(col.sum())


def test_fold_unaligned_pointer(ctx, oracle):
    col = rng_col(ob.I64, 100_001, seed=3, null_frac=0.01)
    d = dev(col)
    got = ctx.fold(capi.F_ALL, ob.I64, d[1:], col.shape[0] - 1)      # base is 8- but not 16-byte aligned
    check_fold(got, oracle_folds(oracle, ob.I64, col[1:]), ob.I64, col.shape[0] - 1)
    col8 = rng_col(ob.U8, 70_001, seed=4)
    got = ctx.fold(capi.F_ALL, ob.U8, dev(col8)[3:], col8.shape[0] - 3)
    check_fold(got, oracle_folds(oracle, ob.U8, col8[3:]), ob.U8, col8.shape[0] - 3)


def test_fold_type_errors(ctx):
    d = dev(np.zeros(16, np.uint8))
    with pytest.raises(capi.RfbError) as e:
        ctx.fold(capi.F_SUM, ob.B8, d, 16)
    assert e.value.kind == "type"


# ---------------------------------------------------------------- fused filter + fold

def unfused(O, op, pt, pred, k, vt, val):
    ids = O.where(O.cmp(op, pt, pred, pt, k))
    return ids, O.at_ids(vt, val, ids)


@pytest.mark.parametrize("op", CMPS)
@pytest.mark.parametrize("n", [0, 1, 17, 16385, 300_007])
def test_filter_fold_same_column_i64(ctx, oracle, op, n):
    col = rng_col(ob.I64, n, seed=11 + n, null_frac=0.02, lo=-1000, hi=1000)
    k = 37
    ids, sel = unfused(oracle, op, ob.I64, col, k, ob.I64, col)
    d = dev(col) if n else None
    got = ctx.filter_fold(op, ob.I64, d, k, capi.F_ALL, ob.I64, d, n)
    check_fold(got, oracle_folds(oracle, ob.I64, sel), ob.I64, ids.shape[0])


@pytest.mark.parametrize("k", [ob.NULL_I64, ob.NULL_I64 + 1, -1, 0, ob.INF_I64])
@pytest.mark.parametrize("op", CMPS)
def test_filter_fold_extreme_constants(ctx, oracle, op, k):
    # i64 comparisons ignore nullness: NULL_I64 is simply the smallest value
    col = np.array([ob.NULL_I64, ob.NULL_I64 + 1, -5, -1, 0, 1, 5, ob.INF_I64 - 1, ob.INF_I64] * 41, np.int64)
    ids, sel = unfused(oracle, op, ob.I64, col, k, ob.I64, col)
    d = dev(col)
    got = ctx.filter_fold(op, ob.I64, d, k, capi.F_ALL, ob.I64, d, col.shape[0])
    check_fold(got, oracle_folds(oracle, ob.I64, sel), ob.I64, ids.shape[0])


@pytest.mark.parametrize("t", [ob.I64, ob.I32])
@pytest.mark.parametrize("k", ["null", "null+1", -3, 0, 7, "max"])
@pytest.mark.parametrize("op", CMPS)
def test_filter_fold_sum_count_fast_path_never_sees_nulls(ctx, oracle, op, k, t):
    """sum/count on the predicate's own column takes the kernel whose predicate was adjusted on the host to reject the
    column's null sentinel (no per-row null test): sum and non-null count must still be the reference's, for every
    operator and for constants at the edges of the domain; `rows` is exact again as soon as RFB_F_ROWS is asked for"""
    info = np.iinfo(ob.NP_OF[t])
    kk = {"null": info.min, "null+1": info.min + 1, "max": info.max}.get(k, k)
    n = 150_001
    col = rng_col(t, n, seed=17, null_frac=0.05, lo=-10, hi=10)
    col[::97] = info.min + 1
    col[1::97] = info.max
    ids, sel = unfused(oracle, op, t, col, kk, t, col)
    d = dev(col)
    got = ctx.filter_fold(op, t, d, kk, capi.F_SUM | capi.F_CNT, t, d, n)
    assert got.sum == int(oracle.fold(ob.SUM, t, sel)[0]) and got.nonnull == int(oracle.fold(ob.CNT, t, sel)[0])
    assert got.rows in (-1, ids.shape[0])
    got = ctx.filter_fold(op, t, d, kk, capi.F_SUM | capi.F_CNT | capi.F_ROWS, t, d, n)
    assert (got.rows, got.sum, got.nonnull) == (ids.shape[0], int(oracle.fold(ob.SUM, t, sel)[0]), int(oracle.fold(ob.CNT, t, sel)[0]))


@pytest.mark.parametrize("pt,vt", [(ob.I64, ob.I64), (ob.I32, ob.I64), (ob.I64, ob.I32), (ob.I32, ob.I32),
                                   (ob.F64, ob.I64), (ob.I64, ob.F64), (ob.F64, ob.F64), (ob.I32, ob.F64)])
@pytest.mark.parametrize("op", [ob.LT, ob.GE, ob.EQ])
def test_filter_fold_two_columns(ctx, oracle, pt, vt, op):
    n = 200_003
    pred = rng_col(pt, n, seed=5, null_frac=0.02, lo=-50, hi=50)
    val = rng_col(vt, n, seed=6, null_frac=0.02)
    if pt == ob.F64:
        pred = np.round(pred)          # make EQ hit
    k = 3
    ids, sel = unfused(oracle, op, pt, pred, k, vt, val)
    got = ctx.filter_fold(op, pt, dev(pred), k, capi.F_ALL, vt, dev(val), n)
    if vt == ob.F64:
        assert got.rows == ids.shape[0] and got.nonnull == int(np.count_nonzero(~np.isnan(sel)))
        assert f64_sum_ok(got.sum, float(oracle.fold(ob.SUM, ob.F64, sel)[0]), oracle.sum_f64_exact(sel))
        assert got.min == float(oracle.fold(ob.MIN, ob.F64, sel)[0]) and got.max == float(oracle.fold(ob.MAX, ob.F64, sel)[0])
    else:
        check_fold(got, oracle_folds(oracle, vt, sel), vt, ids.shape[0])


@pytest.mark.parametrize("op", CMPS)
def test_filter_fold_f64_nan_ordering(ctx, oracle, op):
    # doubles order NaN below everything and NaN == NaN (reference core/ops.h:97,105); -0.0 == +0.0
    col = np.array([np.nan, -np.inf, -1.5, -0.0, 0.0, 1.5, np.inf, np.nan] * 33, np.float64)
    for k in (np.nan, 0.0, -0.0, 1.5, -np.inf, np.inf):
        ids, sel = unfused(oracle, op, ob.F64, col, k, ob.F64, col)
        d = dev(col)
        got = ctx.filter_fold(op, ob.F64, d, k, capi.F_SUM | capi.F_CNT, ob.F64, d, col.shape[0])
        assert got.rows == ids.shape[0], (op, k)
        assert got.nonnull == int(np.count_nonzero(~np.isnan(sel)))


def test_filter_fold_reference_golden_25001_rows(ctx):
    # tests/lang.c:2893-2897 style: > 16384 rows so the reference takes its parallel path
    n = 25001
    col = np.arange(n, dtype=np.int64)
    d = dev(col)
    got = ctx.filter_fold(ob.LT, ob.I64, d, 500, capi.F_SUM | capi.F_CNT, ob.I64, d, n)
    assert (got.nonnull, got.sum) == (500, 124750)
    got = ctx.filter_fold(ob.LT, ob.I64, d, 500, capi.F_SUM | capi.F_CNT | capi.F_ROWS, ob.I64, d, n)
    assert (got.rows, got.nonnull, got.sum) == (500, 500, 124750)


# ---------------------------------------------------------------- compound predicates (and / or) fused with the fold

@pytest.mark.parametrize("npred", [1, 2, 3, 4])
@pytest.mark.parametrize("conj", [True, False])
@pytest.mark.parametrize("vt", [ob.I64, ob.F64])
@pytest.mark.parametrize("n", [1, 1000, 300_007])
def test_multi_filter_fold(ctx, oracle, npred, conj, vt, n):
    """where: (and|or (cmp p1 k1) ...) + fold == per-conjunct masks combined (reference core/logic.c:34-110), where, gather, fold"""
    r = np.random.default_rng(n + npred)
    types = [ob.I64, ob.F64, ob.I64, ob.TIMESTAMP][:npred]
    opsl = [ob.LT, ob.GE, ob.NE, ob.LE][:npred]
    cols, preds, mask = [], [], None
    for t, op in zip(types, opsl):
        c = rng_col(ob.F64 if t == ob.F64 else ob.I64, n, seed=int(r.integers(1 << 30)), null_frac=0.03, lo=-20, hi=20)
        if t == ob.F64:
            c = np.round(c)
        k = 3
        m = oracle.cmp(op, t, c, t, k) != 0
        mask = m if mask is None else ((mask & m) if conj else (mask | m))
        cols.append(c)
        preds.append((op, t, dev(c), k))
    val = rng_col(vt, n, seed=99, null_frac=0.02)
    ids = oracle.where(mask.astype(np.uint8))
    sel = oracle.at_ids(vt, val, ids)
    got = ctx.multi_filter_fold(preds, conj, capi.F_ALL, vt, dev(val), n)
    if vt == ob.F64:
        assert got.rows == ids.shape[0] and got.nonnull == int(np.count_nonzero(~np.isnan(sel)))
        assert f64_sum_ok(got.sum, float(oracle.fold(ob.SUM, ob.F64, sel)[0]), oracle.sum_f64_exact(sel))
        if got.nonnull:
            assert got.min == float(oracle.fold(ob.MIN, ob.F64, sel)[0]) and got.max == float(oracle.fold(ob.MAX, ob.F64, sel)[0])
    else:
        check_fold(got, oracle_folds(oracle, vt, sel), vt, ids.shape[0])


def test_multi_filter_fold_value_column_is_a_predicate_column(ctx, oracle):
    n = 200_001
    x = rng_col(ob.I64, n, seed=5, null_frac=0.02, lo=-100, hi=100)
    d = dev(x)
    got = ctx.multi_filter_fold([(ob.GE, ob.I64, d, -10), (ob.LT, ob.I64, d, 25)], True, capi.F_ALL, ob.I64, d, n)
    sel = x[(x >= -10) & (x < 25)]
    check_fold(got, oracle_folds(oracle, ob.I64, sel), ob.I64, sel.shape[0])
    with pytest.raises(capi.RfbError) as e:
        ctx.multi_filter_fold([(ob.LT, ob.I32, dev(np.zeros(4, np.int32)), 1)], True, capi.F_SUM, ob.I64, dev(np.zeros(4, np.int64)), 4)
    assert e.value.kind == "type"


# ---------------------------------------------------------------- fused (fold (+ (* a b) c))

@pytest.mark.parametrize("n", [0, 1, 3, 16385, 500_001])
def test_fma_fold(ctx, oracle, n):
    a, b, c = (rng_col(ob.F64, n, seed=s, null_frac=0.01, lo=0, hi=1) for s in (1, 2, 3))
    t = oracle.binop(ob.ADD, ob.F64, oracle.binop(ob.MUL, ob.F64, a, ob.F64, b)[0], ob.F64, c)[0] if n else np.empty(0)
    got = ctx.fma_fold(capi.F_ALL, dev(a) if n else None, dev(b) if n else None, dev(c) if n else None, n)
    assert got.rows == n and got.nonnull == int(np.count_nonzero(~np.isnan(t)))
    assert f64_sum_ok(got.sum, float(oracle.fold(ob.SUM, ob.F64, t)[0]), oracle.sum_f64_exact(t))
    if got.nonnull:
        assert got.min == float(oracle.fold(ob.MIN, ob.F64, t)[0]) and got.max == float(oracle.fold(ob.MAX, ob.F64, t)[0])
        avg = float(oracle.fold(ob.AVG, ob.F64, t)[0])
        assert abs(got.avg - avg) <= 2 * abs(avg) * 2.3e-16 + abs(got.sum - float(oracle.fold(ob.SUM, ob.F64, t)[0])) / got.nonnull


def test_fma_fold_exact_integers_bit_exact(ctx, oracle):
    r = np.random.default_rng(9)
    n = 1_000_001
    a, b, c = (r.integers(0, 1 << 10, n).astype(np.float64) for _ in range(3))
    got = ctx.fma_fold(capi.F_SUM | capi.F_CNT, dev(a), dev(b), dev(c), n)
    assert got.sum == float((a * b + c).sum()) and got.nonnull == n


# ---------------------------------------------------------------- fold through a selection vector (MAPFILTER)

@pytest.mark.parametrize("t", [ob.I32, ob.I64, ob.F64, ob.I16, ob.U8])
def test_gather_fold(ctx, oracle, t):
    n = 300_001
    col = rng_col(t, n, seed=21, null_frac=0.03)
    ids = np.sort(np.random.default_rng(22).choice(n, 70_001, replace=False)).astype(np.int64)
    sel = oracle.at_ids(t, col, ids)
    got = ctx.gather_fold(capi.F_ALL, t, dev(col), dev(ids), ids.shape[0])
    if t == ob.F64:
        assert got.rows == ids.shape[0]
        assert f64_sum_ok(got.sum, float(oracle.fold(ob.SUM, ob.F64, sel)[0]), oracle.sum_f64_exact(sel))
    else:
        check_fold(got, oracle_folds(oracle, t, sel), t, ids.shape[0])


# ---------------------------------------------------------------- host layer (host column in, host result out)

@pytest.mark.parametrize("chunk", [0, 1 << 12, 100_000])
def test_filter_fold_host_matches_device_layer(ctx, oracle, chunk):
    n = 1_000_003
    col = rng_col(ob.I64, n, seed=31, null_frac=0.01)
    ids, sel = unfused(oracle, ob.LT, ob.I64, col, 12345, ob.I64, col)
    got, nbytes = ctx.filter_fold_host(ob.LT, ob.I64, col, 12345, capi.F_ALL, ob.I64, col, chunk_rows=chunk)
    check_fold(got, oracle_folds(oracle, ob.I64, sel), ob.I64, ids.shape[0])
    assert nbytes == n * 8
    val = rng_col(ob.F64, n, seed=32, null_frac=0.01)
    ids, sel = unfused(oracle, ob.GE, ob.I64, col, 0, ob.F64, val)
    got, nbytes = ctx.filter_fold_host(ob.GE, ob.I64, col, 0, capi.F_ALL, ob.F64, val, chunk_rows=chunk)
    assert got.rows == ids.shape[0] and nbytes == n * 16
    assert f64_sum_ok(got.sum, float(oracle.fold(ob.SUM, ob.F64, sel)[0]), oracle.sum_f64_exact(sel))
    assert got.min == float(oracle.fold(ob.MIN, ob.F64, sel)[0]) and got.max == float(oracle.fold(ob.MAX, ob.F64, sel)[0])


def test_host_layer_pageable_column_goes_through_the_pinned_ring(ctx, oracle):
    """pageable host columns >= 4 MiB are moved by copier threads through a ring of pinned 16 MiB buffers
    (rfb_copy_h2d); 10M rows = 80 MB = five ring buffers, reused once"""
    n = 10_000_019
    col = rng_col(ob.I64, n, seed=77, null_frac=0.01, lo=-(1 << 30), hi=1 << 30)
    got, nbytes = ctx.filter_fold_host(ob.GE, ob.I64, col, -5, capi.F_ALL, ob.I64, col)
    sel = col[(col >= -5)]
    nn = sel[sel != ob.NULL_I64]
    assert nbytes == 8 * n and got.rows == sel.shape[0] and got.nonnull == nn.shape[0]
    assert got.sum == int(nn.sum(dtype=np.int64)) and got.min == int(nn.min()) and got.max == int(nn.max())
    got2, _ = ctx.fold_host(capi.F_SUM | capi.F_CNT, ob.I64, col, chunk_rows=3_000_000)
    allnn = col[col != ob.NULL_I64]
    assert got2.sum == int(allnn.sum(dtype=np.int64)) and got2.nonnull == allnn.shape[0]


def test_fold_host_empty_and_small(ctx, oracle):
    got, nbytes = ctx.fold_host(capi.F_ALL, ob.I64, np.empty(0, np.int64))
    assert (got.rows, got.sum, got.min, nbytes) == (0, 0, ob.NULL_I64, 0)
    col = rng_col(ob.I32, 77, seed=1, null_frac=0.1)
    got, _ = ctx.fold_host(capi.F_ALL, ob.I32, col)
    check_fold(got, oracle_folds(oracle, ob.I32, col), ob.I32, 77)
