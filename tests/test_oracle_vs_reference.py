"""Differential pin of the CPU oracle (oracle/rf_oracle.c) against the UNMODIFIED reference compiled from its own sources
(oracle/_ref/librayforce_ref.so, built by oracle/Makefile; skipped where that library does not exist).  Seeded random
columns at sizes straddling the reference's 16384-row parallel threshold go through the reference's exported operator
functions (ray_lt, ray_where, filter_collect, ray_sum, index_group, aggr_sum, ray_sort_asc, ...) and through the
oracle; integer results must be bit-identical, fp64 sums within the reduction tolerance.  CPU only."""
import numpy as np
import pytest

from oracle import bindings as ob
from tests.util import rng_col, same_f64, f64_sum_ok, ALL_ARITH_TYPES, typed_col

SIZES = [0, 1, 5, 16383, 16384, 16385, 100_003]
CMPS = [ob.EQ, ob.NE, ob.LT, ob.GT, ob.LE, ob.GE]
ARITH = [ob.ADD, ob.SUB, ob.MUL, ob.DIV, ob.FDIV, ob.MOD, ob.XBAR]
NUM = [ob.I32, ob.I64, ob.F64]


@pytest.mark.parametrize("op", CMPS)
@pytest.mark.parametrize("xt,yt", [(a, b) for a in NUM for b in NUM] + [(ob.I16, ob.I16), (ob.I16, ob.I64), (ob.I32, ob.I16), (ob.F64, ob.I16), (ob.TIMESTAMP, ob.TIMESTAMP), (ob.DATE, ob.DATE)])
def test_cmp(oracle, reference, op, xt, yt):
    n = 20_011
    x = rng_col(xt, n, 1, null_frac=0.05, lo=-9 if xt != ob.U8 else 0, hi=9)
    y = rng_col(yt, n, 2, null_frac=0.05, lo=-9 if yt != ob.U8 else 0, hi=9)
    if xt == ob.F64:
        x = np.round(x)
    if yt == ob.F64:
        y = np.round(y)
    assert np.array_equal(oracle.cmp(op, xt, x, yt, y), reference.cmp(op, xt, x, yt, y))
    if yt in NUM:
        k = y[3]
        assert np.array_equal(oracle.cmp(op, xt, x, yt, k), reference.cmp(op, xt, x, yt, k))
    if xt in NUM:
        k = x[3]
        assert np.array_equal(oracle.cmp(op, xt, k, yt, y), reference.cmp(op, xt, k, yt, y))


@pytest.mark.parametrize("op", CMPS)
def test_cmp_date_vs_timestamp(oracle, reference, op):
    n = 20_011
    r = np.random.default_rng(op)
    d = r.integers(8000, 8010, n).astype(np.int32)
    d[::19] = ob.NULL_I32
    ts = (r.integers(8000, 8010, n) * 86400_000_000_000 + r.integers(-1, 2, n) * 3600_000_000_000).astype(np.int64)
    ts[::23] = ob.NULL_I64
    assert np.array_equal(oracle.cmp(op, ob.DATE, d, ob.TIMESTAMP, ts), reference.cmp(op, ob.DATE, d, ob.TIMESTAMP, ts))
    assert np.array_equal(oracle.cmp(op, ob.TIMESTAMP, ts, ob.DATE, d), reference.cmp(op, ob.TIMESTAMP, ts, ob.DATE, d))


@pytest.mark.parametrize("xt,yt", [(ob.U8, ob.U8), (ob.B8, ob.B8), (ob.U8, ob.I64), (ob.DATE, ob.I32), (ob.I64, ob.TIMESTAMP)])
def test_cmp_type_errors(oracle, reference, xt, yt):
    x, y = rng_col(xt, 50, 1, lo=0, hi=9), rng_col(yt, 50, 2, lo=0, hi=9)
    with pytest.raises(ob.OracleError):
        oracle.cmp(ob.EQ, xt, x, yt, y)
    with pytest.raises(ob.RefError):
        reference.cmp(ob.EQ, xt, x, yt, y)


@pytest.mark.parametrize("n", SIZES[1:])
def test_where_and_gather(oracle, reference, n):
    mask = (np.random.default_rng(n).random(n) < 0.4).astype(np.uint8)
    ids = oracle.where(mask)
    assert np.array_equal(ids, reference.where(mask))
    if ids.shape[0]:
        for t in (ob.I64, ob.F64, ob.I32, ob.I16, ob.U8):
            col = rng_col(t, n, n + t, null_frac=0.02)
            a, b = oracle.at_ids(t, col, ids), reference.at_ids(t, col, ids)
            assert same_f64(a, b) if t == ob.F64 else np.array_equal(a, b)


@pytest.mark.parametrize("t", [ob.U8, ob.I16, ob.I32, ob.I64, ob.TIME])
@pytest.mark.parametrize("op", [ob.SUM, ob.MIN, ob.MAX, ob.AVG])
@pytest.mark.parametrize("n", SIZES[1:])
def test_fold_int(oracle, reference, t, op, n):
    if t == ob.TIME and op == ob.AVG:
        pytest.skip("avg of TIME: not modelled")
    col = rng_col(t, n, n + t, null_frac=0.05)
    a, at = oracle.fold(op, t, col)
    b, bt = reference.fold(op, t, col)
    assert at == bt
    assert same_f64([a], [b]) if at == ob.F64 else int(a) == int(b)


@pytest.mark.parametrize("n", SIZES[1:])
def test_fold_f64(oracle, reference, n):
    col = rng_col(ob.F64, n, n, null_frac=0.05)
    for op in (ob.MIN, ob.MAX):
        assert same_f64([oracle.fold(op, ob.F64, col)[0]], [reference.fold(op, ob.F64, col)[0]])
    exact = oracle.sum_f64_exact(col)
    o, r = float(oracle.fold(ob.SUM, ob.F64, col)[0]), float(reference.fold(ob.SUM, ob.F64, col)[0])
    # both are plain fp64 sums in different orders: each within n ulps of the exact sum of magnitudes
    bound = n * 2.3e-16 * float(np.nansum(np.abs(col)))
    assert abs(o - exact) <= bound and abs(r - exact) <= bound
    ints = np.round(col)
    assert float(oracle.fold(ob.SUM, ob.F64, ints)[0]) == float(reference.fold(ob.SUM, ob.F64, ints)[0])   # exact case: bit-identical


@pytest.mark.parametrize("t", [ob.I64, ob.I32, ob.F64])
def test_fold_all_null_and_empty(oracle, reference, t):
    null = np.nan if t == ob.F64 else np.iinfo(ob.NP_OF[t]).min
    col = np.full(100, null, ob.NP_OF[t])
    for op in (ob.SUM, ob.MIN, ob.MAX, ob.AVG):
        a, at = oracle.fold(op, t, col)
        b, bt = reference.fold(op, t, col)
        assert at == bt and (same_f64([a], [b]) if at == ob.F64 else int(a) == int(b)), (op, a, b)


@pytest.mark.parametrize("n", [1, 16385, 100_003])
def test_filter_fold_pipeline(oracle, reference, n):
    """the unfused pipeline of `select {(sum x) from t where (< x k)}` end to end"""
    col = rng_col(ob.I64, n, n, null_frac=0.02, lo=-1000, hi=1000)
    for op, fold in ((ob.LT, ob.SUM), (ob.GE, ob.MIN), (ob.NE, ob.MAX)):
        ids = oracle.where(oracle.cmp(op, ob.I64, col, ob.I64, 7))
        want = oracle.fold(fold, ob.I64, oracle.at_ids(ob.I64, col, ids))
        got = reference.filter_fold(op, ob.I64, col, 7, fold)
        assert int(want[0]) == int(got[0]) and want[1] == got[1]


def small(t, n, seed):
    a = rng_col(t, n, seed, null_frac=0.04, lo=-50, hi=50)
    if t == ob.F64:
        a = np.round(a * 4) / 4
    a[::17] = 0
    return a


@pytest.mark.parametrize("op", ARITH)
@pytest.mark.parametrize("xt,yt", [(a, b) for a in NUM for b in NUM])
def test_binop(oracle, reference, op, xt, yt):
    n = 20_011
    x, y = small(xt, n, 1), small(yt, n, 2)
    forms = [(x, y), (x, y[5]), (x[5], y), (x, y[0]), (x[0], y)]
    for a, b in forms:
        if xt == ob.I64 and yt == ob.I32 and np.ndim(b) == 0 and np.ndim(a) == 1:
            continue   # reference quirk Q6 (core/math.c:389-391): reads the 8-byte payload of a 4-byte atom; undefined
        (o, ot), (r, rt) = oracle.binop(op, xt, a, yt, b), reference.binop(op, xt, a, yt, b)
        assert ot == rt, (op, xt, yt)
        if ot == ob.F64:
            assert same_f64(o, r, zero_sign=False, max_ulp=1 if op in (ob.FDIV, ob.DIV, ob.MOD) else 0), (op, xt, yt)
        else:
            assert np.array_equal(o, r), (op, xt, yt, np.ndim(a), np.ndim(b))


@pytest.mark.parametrize("op", ARITH)
@pytest.mark.parametrize("xt", ALL_ARITH_TYPES)
def test_binop_type_matrix(oracle, reference, op, xt):
    """Every (operator, operand types, form) case of core/math.c:251-1782 outside I32/I64/F64 x I32/I64/F64: the oracle's matrix
    restatement (oracle/binop_matrix.inc) against the compiled reference — same result type, same bytes, and a type error exactly
    where the reference has no case."""
    n = 4_099
    for yt in ALL_ARITH_TYPES:
        if xt in NUM and yt in NUM:
            continue
        x, y = typed_col(xt, n, 11), typed_col(yt, n, 12)
        for form, (a, b) in ((0, (x, y)), (1, (x, y[5])), (1, (x, y[0])), (1, (x, y[3])), (2, (x[5], y)), (2, (x[0], y)), (2, (x[3], y))):
            want_t = oracle.binop_form(op, form, xt, yt)
            try:
                r, rt = reference.binop(op, xt, a, yt, b)
            except Exception:
                r, rt = None, -1
            if want_t < 0:
                # no kernel case in the reference: it raises a type error too.  One exception (core/math.c:996-997): B8 * I64
                # writes 1-byte results into the 8-byte vector binop_map allocated — no defined result, left out of the matrix
                assert r is None or (op, xt, yt) == (ob.MUL, ob.B8, ob.I64), (op, form, xt, yt)
                continue
            o, ot = oracle.binop(op, xt, a, yt, b)
            assert r is not None, (op, form, xt, yt)
            assert ot == rt == want_t, (op, form, xt, yt, ot, rt)
            if ot == ob.F64:
                assert same_f64(o, r, zero_sign=False, max_ulp=1 if op in (ob.FDIV, ob.DIV, ob.MOD) else 0), (op, form, xt, yt)
            else:
                assert np.array_equal(o, r), (op, form, xt, yt, o[:20], r[:20], np.flatnonzero(o != r)[:10])


def test_binop_matrix_agrees_with_plain_numeric_rules(oracle):
    """inside I32/I64/F64 the generated matrix and the hand-written typing rules (pinned in test_binop) must say the same"""
    for op in ARITH:
        for xt in NUM:
            for yt in NUM:
                for form in (0, 1, 2):
                    assert oracle.binop_form(op, form, xt, yt) == oracle.binop_type(op, xt, yt), (op, form, xt, yt)


@pytest.mark.parametrize("op", [ob.ROUND, ob.FLOOR, ob.CEIL])
def test_unop(oracle, reference, op):
    x = np.concatenate([rng_col(ob.F64, 20_011, op, null_frac=0.02, lo=-1e6, hi=1e6), np.array([0.5, -0.5, 1.5, 2.5, -1.5, 4.0, -4.0, 0.0])])
    assert same_f64(oracle.unop_f64(op, x), reference.unop_f64(op, x), zero_sign=False)


A_OPS = [ob.SUM, ob.MIN, ob.MAX, ob.COUNT, ob.AVG]


def reference_scope_is_safe(n, cores):
    """Reference bug found while pinning (Q12 in DESIGN.md): index_scope_i64 (core/index.c:402-435) hands every worker but
    the last a page-aligned chunk (pool_chunk_aligned) and gives the last one `len - (chunks-1)*chunk` rows; when
    (chunks-1)*chunk exceeds len (e.g. len = 16385 on 8 threads: 7*2560 > 16385) workers read past the end of the key
    column (confirmed with an ASan build: heap-buffer-overflow at core/index.c:391) and the group index is garbage.
    Differential group-by cases are only run at lengths where the reference stays in bounds."""
    if n < 16384:
        return True
    chunk = -(-(-(-n // cores)) // 512) * 512
    return (cores - 1) * chunk < n


@pytest.mark.parametrize("n,card", [(7, 3), (16000, 100), (60_000, 100), (100_003, 1000), (300_007, 70_000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_group_dense_and_aggregates(oracle, reference, n, card, filtered):
    r = np.random.default_rng(n + card)
    keys = (r.integers(0, card, n) - 500).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    if not reference_scope_is_safe(n if filt is None else filt.shape[0], reference.cores):
        pytest.skip("reference index_scope_i64 reads out of bounds at this length / thread count (Q12)")
    gids, firsts, info = oracle.group_i64(keys, filt)
    for vt in (ob.I64, ob.F64, ob.I32, ob.I16, ob.TIME, ob.TIMESTAMP):
        val = rng_col(vt, n, vt, null_frac=0.001, lo=-1000, hi=1000)
        if vt == ob.F64:
            val = np.round(val * 8) / 8
        # the reference's non-parted drivers (core/aggr.c:1107-1150, 1152-1315, 1380-1453, 2013-2133)
        ops = {ob.I64: A_OPS, ob.F64: A_OPS, ob.I32: [ob.COUNT, ob.AVG], ob.I16: [ob.SUM, ob.MIN, ob.MAX, ob.AVG],
               ob.TIME: [ob.MIN, ob.MAX, ob.COUNT, ob.AVG], ob.TIMESTAMP: [ob.MIN, ob.MAX, ob.COUNT]}[vt]
        ref = reference.group_aggr(keys, vt, val, ops, filt)
        assert ref["groups"] == info.groups and ref["index_type"] == info.index_type
        if ref["first_ids"] is not None:
            assert np.array_equal(ref["first_ids"], firsts)
        for op in ops:
            want, wt = oracle.aggr(op, vt, val, gids, info.groups, filt)
            got, gt = ref["results"][op]
            assert gt == wt, (op, vt)
            if not info.dense:   # hash path: the reference's group ORDER depends on its thread count
                want, got = np.sort(want), np.sort(got)
            assert same_f64(want, got, zero_sign=False) if wt == ob.F64 else np.array_equal(want, got), (op, vt)


@pytest.mark.parametrize("op,vt", [(ob.SUM, ob.I32), (ob.SUM, ob.TIME), (ob.SUM, ob.TIMESTAMP), (ob.MIN, ob.I32), (ob.MAX, ob.I32),
                                   (ob.AVG, ob.TIMESTAMP), (ob.COUNT, ob.I16), (ob.COUNT, ob.U8)])
def test_grouped_aggregate_type_errors(oracle, reference, op, vt):
    keys = np.arange(50, dtype=np.int64) % 5
    val = rng_col(vt, 50, 1, lo=0, hi=9)
    gids, firsts, info = oracle.group_i64(keys)
    with pytest.raises(ob.OracleError):
        oracle.aggr(op, vt, val, gids, info.groups)
    with pytest.raises(ob.RefError):
        reference.group_aggr(keys, vt, val, [op])


@pytest.mark.parametrize("n", [1000, 60_000])
def test_multi_key_group_by_through_rayfall_select(oracle, reference, n):
    """(select {s: (sum v) c: (count v) from: t by: {a: a b: b c: c}}) through the reference's evaluator: its key-tuple order
    is first occurrence, and sums/counts per tuple match the oracle's multi-key index + grouped aggregates"""
    import ctypes as C
    if not reference_scope_is_safe(n, reference.cores):
        pytest.skip("reference index_scope_i64 reads out of bounds at this length / thread count (Q12)")
    r = np.random.default_rng(n)
    a, b, c = r.integers(0, 7, n), r.integers(-3, 4, n) * 1000, r.integers(100, 104, n)
    v = r.integers(-50, 50, n)
    for name, arr in (("mk_a", a), ("mk_b", b), ("mk_c", c), ("mk_v", v)):
        o = reference.eval("(set %s (til %d))" % (name, n))
        np.frombuffer((C.c_char * (n * 8)).from_address(o + 16), dtype=np.int64)[:] = arr
    reference.eval("(set mk_t (table [a b c v] (list mk_a mk_b mk_c mk_v)))")
    res = reference.eval("(select {s: (sum v) n: (count v) from: mk_t by: {a: a b: b c: c}})")
    cols = [reference.to_numpy(x, drop=False)[0] for x in reference.list_items(reference.list_items(res)[1])]
    gids, firsts, groups = oracle.group_multi([a, b, c])
    assert groups == cols[0].shape[0]
    want = np.stack([a[firsts], b[firsts], c[firsts], oracle.aggr(ob.SUM, ob.I64, v, gids, groups)[0],
                     oracle.aggr(ob.COUNT, ob.I64, v, gids, groups)[0]])
    got = np.stack(cols)
    if n >= 16384:
        # above its parallel threshold the reference's group ORDER depends on its thread count (the same Q9 as the
        # single-key hash path): compare per key tuple
        want = want[:, np.lexsort(want[:3][::-1])]
        got = got[:, np.lexsort(got[:3][::-1])]
    assert np.array_equal(got, want)


def test_group_sparse_as_key_sorted_sets(oracle, reference):
    """hash path (range > len): the reference's group order depends on its thread count, so compare per key"""
    n = 100_003
    if not reference_scope_is_safe(n, reference.cores):
        pytest.skip("reference index_scope_i64 reads out of bounds at this length / thread count (Q12)")
    r = np.random.default_rng(4)
    pool = r.integers(-(1 << 60), 1 << 60, 5000).astype(np.int64)
    keys = pool[r.integers(0, 5000, n)]
    val = r.integers(-100, 100, n).astype(np.int64)
    gids, firsts, info = oracle.group_i64(keys)
    assert info.dense == 0
    ref = reference.group_aggr(keys, ob.I64, val, [ob.SUM, ob.COUNT])
    assert ref["groups"] == info.groups
    okeys = keys[firsts]
    osum, ocnt = oracle.aggr(ob.SUM, ob.I64, val, gids, info.groups)[0], oracle.aggr(ob.COUNT, ob.I64, val, gids, info.groups)[0]
    want = dict(zip(okeys.tolist(), zip(osum.tolist(), ocnt.tolist())))
    # reference group keys: first row of each group is not exposed on this path; rebuild per-key totals from numpy instead
    uk, inv = np.unique(keys, return_inverse=True)
    rs = np.zeros(uk.shape[0], np.int64)
    np.add.at(rs, inv, val)
    rc = np.bincount(inv)
    assert want == dict(zip(uk.tolist(), zip(rs.tolist(), rc.tolist())))
    assert sorted(ref["results"][ob.SUM][0].tolist()) == sorted(osum.tolist())
    assert sorted(ref["results"][ob.COUNT][0].tolist()) == sorted(ocnt.tolist())


@pytest.mark.parametrize("t", [ob.U8, ob.I16, ob.I32, ob.I64, ob.F64])
@pytest.mark.parametrize("desc", [False, True])
@pytest.mark.parametrize("n", [1, 100, 70_001])
def test_sort(oracle, reference, t, desc, n):
    col = rng_col(t, n, n + t, null_frac=0.05, lo=-40 if t != ob.U8 else 0, hi=40)
    if t == ob.F64:
        col = np.round(col)
        col[::11] = -0.0
    assert np.array_equal(oracle.sort(t, col, desc), reference.sort(t, col, desc))
    wide = rng_col(t, n, n + t + 1, null_frac=0.05)
    assert np.array_equal(oracle.sort(t, wide, desc), reference.sort(t, wide, desc))


# ---------------------------------------------------------------- med / dev / collect / row

def close_f64(a, b, rel=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=rel, atol=0)


@pytest.mark.parametrize("t", [ob.U8, ob.I16, ob.I64])
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 40_001])
@pytest.mark.parametrize("nulls", [False, True])
def test_ungrouped_med(oracle, reference, t, n, nulls):
    """ray_med (core/math.c:2529-2626) incl. its quirk: the middle of the NON-NULL count indexes the column sorted WITH its nulls"""
    col = rng_col(t, n, n + t, null_frac=0.2 if nulls else 0.0, lo=-50 if t != ob.U8 else 0, hi=50)
    a, b = oracle.med(t, col), float(reference.fold(ob.MED, t, col)[0])
    assert same_f64([a], [b])


@pytest.mark.parametrize("t", [ob.I32, ob.F64, ob.DATE])
def test_ungrouped_med_type_errors(oracle, reference, t):
    col = rng_col(t, 10, 1, lo=0, hi=9)
    with pytest.raises(ob.OracleError):
        oracle.med(t, col)
    with pytest.raises(ob.RefError):
        reference.fold(ob.MED, t, col)


@pytest.mark.parametrize("t", [ob.U8, ob.I16, ob.I32, ob.TIME, ob.I64, ob.F64])
@pytest.mark.parametrize("n", [1, 2, 7, 1000, 40_001])
def test_ungrouped_dev(oracle, reference, t, n):
    """ray_dev (core/math.c:2628-2700): two passes; the sums run in a thread-dependent order in the reference, so the results
    agree to rounding (exactly on inputs whose partial sums are exact)"""
    col = rng_col(t, n, n + t, null_frac=0.1, lo=-100 if t != ob.U8 else 0, hi=100)
    a, b = oracle.dev(t, col), float(reference.fold(ob.DEV, t, col)[0])
    assert close_f64([a], [b])
    allnull = np.full(5, np.nan if t == ob.F64 else (0 if t == ob.U8 else np.iinfo(ob.NP_OF[t]).min), ob.NP_OF[t])
    if t != ob.U8:
        assert np.isnan(oracle.dev(t, allnull)) and np.isnan(float(reference.fold(ob.DEV, t, allnull)[0]))


@pytest.mark.parametrize("n,card", [(9, 3), (20_000, 50), (100_003, 3000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_grouped_med_and_dev(oracle, reference, n, card, filtered):
    r = np.random.default_rng(n + card)
    keys = r.integers(0, card, n).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    if not reference_scope_is_safe(n if filt is None else filt.shape[0], reference.cores):
        pytest.skip("reference index_scope_i64 reads out of bounds at this length / thread count (Q12)")
    gids, firsts, info = oracle.group_i64(keys, filt)
    for vt, ops in ((ob.I64, [ob.MED, ob.DEV]), (ob.F64, [ob.MED, ob.DEV]), (ob.TIMESTAMP, [ob.MED, ob.DEV]), (ob.I32, [ob.MED, ob.DEV]),
                    (ob.I16, [ob.DEV]), (ob.TIME, [ob.DEV])):
        val = rng_col(vt, n, vt, null_frac=0.01, lo=-1000, hi=1000)
        if vt == ob.F64:
            val = np.round(val * 8) / 8
        ref = reference.group_aggr(keys, vt, val, ops, filt)
        for op in ops:
            want, wt = oracle.aggr(op, vt, val, gids, info.groups, filt)
            got, gt = ref["results"][op]
            assert gt == wt == ob.F64
            assert same_f64(want, got) if op == ob.MED else close_f64(want, got, 1e-9), (op, vt)


def ref_aggr_chunks(rows, groups, width, cores):
    """pool_split_by_mem (core/pool.c:450-478): how many worker chunks aggr_map cuts `rows` rows into"""
    if rows < 16384 or rows <= cores:
        return 1
    mem = groups * width
    if mem > (64 << 20):
        return 1
    return max(1, min(cores, (64 << 20) // mem)) if mem > 0 else cores


@pytest.mark.parametrize("n,card", [(9, 3), (16000, 100), (20_000, 50), (100_003, 3000), (300_007, 70_000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_grouped_first_and_last(oracle, reference, n, card, filtered):
    """aggr_first = the value at the group's first row (first_ids fast path, nulls included); aggr_last = the last non-null
    value inside the first WORKER CHUNK that has one (AGGR_COLLECT keeps the first non-null partial): above 16384 rows the
    reference's own answer depends on its executor count, which the oracle takes as a parameter"""
    r = np.random.default_rng(n * 7 + card)
    keys = r.integers(0, card, n).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    rows = n if filt is None else filt.shape[0]
    if not reference_scope_is_safe(rows, reference.cores):
        pytest.skip("reference index_scope_i64 reads out of bounds at this length / thread count (Q12)")
    gids, firsts, info = oracle.group_i64(keys, filt)
    for vt in (ob.I64, ob.F64, ob.I32, ob.I16, ob.TIME, ob.TIMESTAMP, ob.DATE, ob.U8):
        val = rng_col(vt, n, vt + 3, null_frac=0.3, lo=-1000 if vt != ob.U8 else 0, hi=1000 if vt != ob.U8 else 200)
        ops = [ob.FIRST] if vt == ob.U8 else [ob.FIRST, ob.LAST]
        ref = reference.group_aggr(keys, vt, val, ops, filt)
        want, wt = oracle.aggr(ob.FIRST, vt, val, gids, info.groups, filt)
        got, gt = ref["results"][ob.FIRST]
        if not info.dense:
            want, got = np.sort(want), np.sort(got)
        assert gt == wt and (same_f64(want, got, zero_sign=False) if wt == ob.F64 else np.array_equal(want, got)), vt
        if vt == ob.U8:
            continue
        chunks = ref_aggr_chunks(rows, info.groups, np.dtype(ob.NP_OF[vt]).itemsize, reference.cores)
        want, wt = oracle.aggr_last(vt, val, gids, info.groups, chunks, filt)
        got, gt = ref["results"][ob.LAST]
        if not info.dense:
            want, got = np.sort(want), np.sort(got)
        assert gt == wt, (vt, gt, wt)
        assert same_f64(want, got, zero_sign=False) if wt == ob.F64 else np.array_equal(want, got), (vt, chunks)


def test_grouped_dev_type_error(oracle, reference):
    keys = np.arange(50, dtype=np.int64) % 5
    val = rng_col(ob.U8, 50, 1, lo=0, hi=9)
    gids, firsts, info = oracle.group_i64(keys)
    with pytest.raises(ob.OracleError):
        oracle.aggr(ob.DEV, ob.U8, val, gids, info.groups)
    with pytest.raises(ob.RefError):
        reference.group_aggr(keys, ob.U8, val, [ob.DEV])


@pytest.mark.parametrize("filtered", [False, True])
def test_group_rows_collect(oracle, reference, filtered):
    """aggr_row / aggr_collect (core/aggr.c:3021-3136): per-group row-id / value lists in push order"""
    n, card = 20_011, 37
    r = np.random.default_rng(3)
    keys = r.integers(0, card, n).astype(np.int64)
    val = rng_col(ob.I64, n, 5, null_frac=0.01)
    filt = np.sort(r.choice(n, n // 2, replace=False)).astype(np.int64) if filtered else None
    if not reference_scope_is_safe(n if filt is None else filt.shape[0], reference.cores):
        pytest.skip("Q12")
    gids, firsts, info = oracle.group_i64(keys, filt)
    rows, offs = oracle.group_rows(gids, info.groups, filt)
    got_rows, got_vals = reference.group_lists(keys, ob.I64, val, filt)
    assert len(got_rows) == info.groups
    for g in range(info.groups):
        assert np.array_equal(got_rows[g], rows[offs[g]:offs[g + 1]])
        assert np.array_equal(got_vals[g], val[rows[offs[g]:offs[g + 1]]])


# ---------------------------------------------------------------- equi-join row matching

@pytest.mark.parametrize("nb,np_,card", [(1, 1, 1), (100, 300, 40), (50_000, 80_000, 20_000), (50_000, 80_000, 10**12)])
def test_find_single_key(oracle, reference, nb, np_, card):
    """ray_find -> index_find_i64 (core/index.c:1507-1574): dense (range <= MAX_RANGE) and hashed key domains.  Keys are kept
    non-negative and non-null: the reference's hash path indexes its table with (i64)key % size (core/hash.c:104)."""
    r = np.random.default_rng(nb + np_)
    pool = r.integers(0, card, max(2, nb // 2)).astype(np.int64)
    build = pool[r.integers(0, pool.shape[0], nb)]
    probe = np.concatenate([pool[r.integers(0, pool.shape[0], np_ // 2)], r.integers(0, card, np_ - np_ // 2)]).astype(np.int64)
    assert np.array_equal(oracle.find_rows([build], [probe]), reference.find(build, probe))


@pytest.mark.parametrize("ncols", [2, 3])
@pytest.mark.parametrize("nb,np_", [(10, 30), (40_000, 70_000)])
def test_join_index_multi_key(oracle, reference, ncols, nb, np_):
    """index_left_join_obj / index_inner_join_obj (core/index.c:2886-3000), the hashed multi-column path"""
    r = np.random.default_rng(nb * ncols)
    bcols = [r.integers(0, 12 + c, nb).astype(np.int64) * (10**9 if c == 0 else 1) for c in range(ncols)]
    pcols = [r.integers(0, 14 + c, np_).astype(np.int64) * (10**9 if c == 0 else 1) for c in range(ncols)]
    bcols[-1][::7] = ob.NULL_I64
    pcols[-1][::5] = ob.NULL_I64                     # nulls are keys like any other (bitwise row compare)
    want = oracle.find_rows(bcols, pcols)
    assert np.array_equal(want, reference.join_index(pcols, bcols))
    pi, bi = oracle.inner_join(bcols, pcols)
    rl, rr = reference.join_index(pcols, bcols, inner=True)
    assert np.array_equal(pi, rl[:pi.shape[0]]) and np.array_equal(bi, rr[:bi.shape[0]])


# ---------------------------------------------------------------- parted aggregates

def test_parted_aggregates_through_rayfall(oracle, reference, tmp_path):
    """PARTED_MAP (core/aggr.c:183-260): a parted table written and re-opened by the reference itself (set-splayed / get-parted,
    as tests/parted.c does), aggregated over all partitions and per partition; the oracle gets the same columns from numpy"""
    root = str(tmp_path) + "/"
    n, days = 2000, 4
    setup = ('(do (set dbpath "%s") (set n %d)'
             ' (set gen (fn [day] (let p (format "%%/%%/a/" dbpath (+ 2024.01.01 day)))'
             '   (let t (table [v f] (list (- (%% (+ (* (til n) 7919) (* day 13)) 1000) 300) (div (as \'F64 (%% (+ (* (til n) 31) day) 640)) 8.0))))'
             '   (set-splayed p t)))'
             ' (map gen (til %d)) (set t (get-parted dbpath \'a)) 0)') % (root, n, days)
    reference.eval(setup)
    til = np.arange(n, dtype=np.int64)
    v = [((til * 7919 + d * 13) % 1000 - 300) for d in range(days)]
    f = [((til * 31 + d) % 640).astype(np.float64) / 8.0 for d in range(days)]

    def ref(expr):
        return reference.to_numpy(reference.eval(expr))[0]

    for name, op in (("sum", ob.SUM), ("min", ob.MIN), ("max", ob.MAX), ("avg", ob.AVG)):
        for col, parts, t in (("v", v, ob.I64), ("f", f, ob.F64)):
            got = ref("(at (select {from: t r: (%s %s)}) 'r)" % (name, col))
            want, wt = oracle.parted_aggr(op, t, parts, True)
            assert same_f64(want, got) if wt == ob.F64 else np.array_equal(want, got), (name, col)
            got = ref("(at (select {from: t by: Date r: (%s %s)}) 'r)" % (name, col))
            want, wt = oracle.parted_aggr(op, t, parts, False)
            assert same_f64(want, got) if wt == ob.F64 else np.array_equal(want, got), (name, col, "by Date")


@pytest.mark.parametrize("nx,ny,card", [(100, 30, 40), (80_000, 50_000, 20_000), (80_000, 50_000, 10**12)])
def test_in_single_key(oracle, reference, nx, ny, card):
    """ray_in -> index_in_i64_i64 (core/index.c:1291-1370) == "ray_find found a row" (non-negative keys: see test_find_single_key)"""
    r = np.random.default_rng(nx + ny)
    pool = r.integers(0, card, max(2, ny // 2)).astype(np.int64)
    y = pool[r.integers(0, pool.shape[0], ny)]
    x = np.concatenate([pool[r.integers(0, pool.shape[0], nx // 2)], r.integers(0, card, nx - nx // 2)]).astype(np.int64)
    want = (oracle.find_rows([y], [x]) != ob.NULL_I64).astype(np.uint8)
    assert np.array_equal(want, reference.isin(x, y).astype(np.uint8))


def asof_tables(ncols, nb, np_, seed, tt=ob.I64):
    """build side sorted by time (asof joins search each key's rows in row order), probe side anywhere in the time range"""
    r = np.random.default_rng(seed)
    bcols = [r.integers(0, 6 + c, nb).astype(np.int64) for c in range(ncols)]
    pcols = [r.integers(0, 7 + c, np_).astype(np.int64) for c in range(ncols)]
    bt = np.sort(r.integers(0, 1_000_000, nb)).astype(ob.NP_OF[tt])
    pt = r.integers(-10, 1_000_100, np_).astype(ob.NP_OF[tt])
    return bcols, bt, pcols, pt


@pytest.mark.parametrize("ncols", [1, 2])
@pytest.mark.parametrize("tt", [ob.I64, ob.I32])
@pytest.mark.parametrize("nb,np_", [(20, 50), (30_000, 50_000)])
def test_asof_join_index(oracle, reference, ncols, tt, nb, np_):
    """index_asof_join_obj (core/index.c:3194-3268)"""
    bcols, bt, pcols, pt = asof_tables(ncols, nb, np_, nb + ncols, tt)
    want = oracle.asof_join(bcols, tt, bt, pcols, pt)
    assert np.array_equal(want, reference.asof_index(pcols, tt, pt, bcols, bt))
    assert (want != ob.NULL_I64).any() and (want == ob.NULL_I64).any()


@pytest.mark.parametrize("n,card", [(5, 5), (6, 4), (100, 37), (1000, 900), (30_000, 5000), (200_003, 150_000)])
def test_distinct_sparse_hash_branch(oracle, reference, n, card):
    """ray_distinct -> index_distinct_i64, hash branch (core/index.c:579-603): range > len and > 2^20 -> an open-addressing table of
    next_prime(ceil(len / 0.75)) slots, rows inserted in row order, the result is the table in SLOT order (non-negative keys)"""
    r = np.random.default_rng(n + card)
    pool = r.integers(0, 1 << 62, card).astype(np.int64)
    keys = pool[r.integers(0, card, n)]
    keys[0], keys[-1] = 0, (1 << 62) + 12345                       # force a huge range
    v = reference.vec(ob.I64, keys)
    got = reference.to_numpy(reference.call1("ray_distinct", v))[0]
    reference.drop(v)
    assert np.array_equal(oracle.distinct(keys), got)


@pytest.mark.parametrize("n,card,kmin", [(1, 1, 5), (1000, 40, -20), (200_003, 5000, -2500), (50_000, 900_000, 17)])
def test_distinct_dense(oracle, reference, n, card, kmin):
    """ray_distinct -> index_distinct_i64, direct-addressing branch (core/index.c:551-577): ascending distinct keys"""
    keys = (np.random.default_rng(n).integers(0, card, n) + kmin).astype(np.int64)
    v = reference.vec(ob.I64, keys)
    got = reference.to_numpy(reference.call1("ray_distinct", v))[0]
    reference.drop(v)
    assert np.array_equal(oracle.distinct(keys), got)


@pytest.mark.parametrize("n", [800, 50_000])
def test_multi_key_group_by_row_hash_path_through_rayfall_select(oracle, reference, n):
    """the same query with key columns whose ranges do not multiply into an i64 (core/index.c:2556-2729, row hashes): the tuples,
    their sums and counts must match the oracle's multi-key index + grouped aggregates (compared per key tuple: the reference's
    group order on this path is its hash tables' order)"""
    import ctypes as C
    if not reference_scope_is_safe(n, reference.cores):
        pytest.skip("Q12")
    r = np.random.default_rng(n + 1)
    pa, pb = r.integers(-(1 << 61), 1 << 61, 9).astype(np.int64), r.integers(-(1 << 61), 1 << 61, 11).astype(np.int64)
    a, b = pa[r.integers(0, 9, n)], pb[r.integers(0, 11, n)]
    v = r.integers(-50, 50, n)
    for name, arr in (("mh_a", a), ("mh_b", b), ("mh_v", v)):
        o = reference.eval("(set %s (til %d))" % (name, n))
        np.frombuffer((C.c_char * (n * 8)).from_address(o + 16), dtype=np.int64)[:] = arr
    reference.eval("(set mh_t (table [a b v] (list mh_a mh_b mh_v)))")
    res = reference.eval("(select {s: (sum v) n: (count v) from: mh_t by: {a: a b: b}})")
    cols = [reference.to_numpy(x, drop=False)[0] for x in reference.list_items(reference.list_items(res)[1])]
    gids, firsts, groups = oracle.group_multi([a, b])
    assert groups == cols[0].shape[0]
    want = np.stack([a[firsts], b[firsts], oracle.aggr(ob.SUM, ob.I64, v, gids, groups)[0], oracle.aggr(ob.COUNT, ob.I64, v, gids, groups)[0]])
    got = np.stack(cols)
    want = want[:, np.lexsort(want[:2][::-1])]
    got = got[:, np.lexsort(got[:2][::-1])]
    assert np.array_equal(got, want)


# ---------------------------------------------------------------- window joins (oracle only so far: DESIGN.md §10)

def _set_col(reference, name, arr, t):
    import ctypes as C
    n = arr.shape[0]
    src = {ob.I64: "(til %d)", ob.TIME: "(as 'TIME (til %d))", ob.F64: "(as 'F64 (til %d))"}[t] % n
    o = reference.eval("(set %s %s)" % (name, src))
    assert reference.type_of(o) == t
    np.frombuffer((C.c_char * (n * np.dtype(ob.NP_OF[t]).itemsize)).from_address(o + 16), dtype=ob.NP_OF[t])[:] = arr


@pytest.mark.parametrize("nl,nr,lkeys,rkeys,span", [(50, 400, 5, 6, 100_000), (2000, 30_000, 40, 35, 1_000_000), (300, 50, 3, 9, 5_000)])
@pytest.mark.parametrize("vt", [ob.I64, ob.F64])
def test_window_join_aggregates_through_rayfall(oracle, reference, nl, nr, lkeys, rkeys, span, vt):
    """(window-join / window-join1 [Sym Time] intervals trades quotes {r: (agg Bid)}): the reference sorts quotes by (Sym, Time),
    keeps the [first, last] row block per key and folds the rows inside every trade's window (core/join.c:358-485,
    core/index.c:3287-3346, core/aggr.c:39-72, 131-160).  Keys missing on either side, windows without rows, nulls in Bid."""
    r = np.random.default_rng(nl + nr)
    ls = r.integers(0, lkeys, nl).astype(np.int64)
    lt = np.sort(r.integers(0, span, nl)).astype(np.int32)
    rs = r.integers(0, rkeys, nr).astype(np.int64)
    rt = r.integers(0, span, nr).astype(np.int32)
    bid = rng_col(vt, nr, 7, null_frac=0.02, lo=-50, hi=50)
    if vt == ob.F64:
        bid = np.round(bid * 4) / 4
    for name, arr, t in (("wj_ls", ls, ob.I64), ("wj_lt", lt, ob.TIME), ("wj_rs", rs, ob.I64), ("wj_rt", rt, ob.TIME), ("wj_bid", bid, vt)):
        _set_col(reference, name, arr, t)
    reference.eval("(set wj_trades (table [Sym Time] (list wj_ls wj_lt)))")
    reference.eval("(set wj_quotes (table [Sym Time Bid] (list wj_rs wj_rt wj_bid)))")
    reference.eval("(set wj_iv (map-left + [-2000 3000] (at wj_trades 'Time)))")
    order = np.lexsort((rt, rs))                       # xasc [Sym Time]: stable
    for fn, jt in (("window-join", 0), ("window-join1", 1)):
        for name, op in (("min", ob.MIN), ("max", ob.MAX), ("sum", ob.SUM), ("count", ob.COUNT), ("avg", ob.AVG)):
            got = reference.to_numpy(reference.eval("(at (%s [Sym Time] wj_iv wj_trades wj_quotes {r: (%s Bid)}) 'r)" % (fn, name)))[0]
            want, wt = oracle.window_aggr(op, vt, bid[order], [rs[order]], rt[order], [ls], lt - 2000, lt + 3000, jt)
            assert same_f64(want, got) if wt == ob.F64 else np.array_equal(want, got), (fn, name)


@pytest.mark.parametrize("seed", range(12))
def test_join_pins_randomized(oracle, reference, seed):
    """randomized shapes for the row-matching pins: 1-3 key columns, tiny to mid sizes, low cardinality (many duplicate keys and
    times), keys missing on either side"""
    r = np.random.default_rng(seed)
    nb, np_ = int(r.integers(1, 3000)), int(r.integers(1, 3000))
    nc, card = int(r.integers(1, 4)), int(r.integers(1, 12))
    b = [r.integers(0, card, nb).astype(np.int64) for _ in range(nc)]
    p = [r.integers(0, card + 2, np_).astype(np.int64) for _ in range(nc)]
    bt, pt = np.sort(r.integers(0, 50, nb)).astype(np.int64), r.integers(-5, 60, np_).astype(np.int64)
    assert np.array_equal(oracle.asof_join(b, ob.I64, bt, p, pt), reference.asof_index(p, ob.I64, pt, b, bt))
    if nc >= 2:
        assert np.array_equal(oracle.find_rows(b, p), reference.join_index(p, b))
    else:
        ids = oracle.find_rows(b, p)
        assert np.array_equal(ids, reference.find(b[0], p[0]))
        assert np.array_equal((ids != ob.NULL_I64).astype(np.uint8), reference.isin(p[0], b[0]).astype(np.uint8))
        v = reference.vec(ob.I64, b[0])
        got = reference.to_numpy(reference.call1("ray_distinct", v))[0]
        reference.drop(v)
        assert np.array_equal(oracle.distinct(b[0]), got)
