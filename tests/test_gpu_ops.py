"""GPU parity for the operator-exact building blocks (k_map.cu, k_select.cu, k_group.cu, k_sort.cu) against the CPU
oracle through the C ABI: comparisons, where, gather, arithmetic, round/floor/ceil, group index, grouped aggregates,
fused group-by and the stable key sort.  Edge cases follow the reference's tests (tests/lang.c test_lang_cmp/
test_lang_math/test_lang_group, tests/sort.c): empty and 1-row inputs, sizes straddling 16384, nulls/NaN, atoms on
either side, length mismatches, unsupported type mixes."""
import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200 import capi
from tests.util import ALL_ARITH_TYPES, dev, host, rng_col, same_f64, typed_col

pytestmark = pytest.mark.gpu

SIZES = [1, 2, 31, 1023, 16385, 200_003]
CMPS = [ob.EQ, ob.NE, ob.LT, ob.GT, ob.LE, ob.GE]
ARITH = [ob.ADD, ob.SUB, ob.MUL, ob.DIV, ob.FDIV, ob.MOD, ob.XBAR]
ALL_T = [ob.U8, ob.I16, ob.I32, ob.I64, ob.F64]
CMP_T = [ob.I16, ob.I32, ob.I64, ob.F64]      # the reference's vector comparison matrix (core/cmp.c:77-258)


# ---------------------------------------------------------------- comparisons -> mask

@pytest.mark.parametrize("op", CMPS)
@pytest.mark.parametrize("xt,yt", [(a, b) for a in CMP_T for b in CMP_T] + [(ob.DATE, ob.DATE), (ob.TIMESTAMP, ob.TIMESTAMP)])
def test_cmp_vector_vector(ctx, oracle, op, xt, yt):
    n = 40_007
    x = rng_col(xt, n, seed=xt * 7 + 1, null_frac=0.05, lo=-20 if xt != ob.U8 else 0, hi=20)
    y = rng_col(yt, n, seed=yt * 11 + 2, null_frac=0.05, lo=-20 if yt != ob.U8 else 0, hi=20)
    if xt == ob.F64:
        x = np.round(x)
    if yt == ob.F64:
        y = np.round(y)
    got = host(ctx.cmp(op, xt, dev(x), yt, dev(y)))
    assert np.array_equal(got, oracle.cmp(op, xt, x, yt, y))


@pytest.mark.parametrize("op", CMPS)
@pytest.mark.parametrize("t,kt,k", [(ob.I64, ob.I64, 3), (ob.I64, ob.I64, ob.NULL_I64), (ob.I32, ob.I64, -2), (ob.I64, ob.I32, 0),
                                    (ob.I64, ob.F64, 2.5), (ob.F64, ob.I64, 1), (ob.F64, ob.F64, float("nan")),
                                    (ob.F64, ob.F64, -0.0), (ob.I16, ob.I16, ob.NULL_I16), (ob.TIME, ob.TIME, 7)])
@pytest.mark.parametrize("n", [1, 17, 70_001])
def test_cmp_vector_atom_both_sides(ctx, oracle, op, t, kt, k, n):
    x = rng_col(t, n, seed=n + t, null_frac=0.05, lo=-9 if t != ob.U8 else 0, hi=9)
    if t == ob.F64:
        x = np.round(x)
        x[::13] = -0.0
    d = dev(x)
    assert np.array_equal(host(ctx.cmp(op, t, d, kt, k)), oracle.cmp(op, t, x, kt, k))
    assert np.array_equal(host(ctx.cmp(op, kt, k, t, d)), oracle.cmp(op, kt, k, t, x))


@pytest.mark.parametrize("xt,yt", [(ob.U8, ob.U8), (ob.B8, ob.B8), (ob.U8, ob.I64), (ob.DATE, ob.I32), (ob.I64, ob.TIMESTAMP)])
def test_cmp_type_errors_match_the_reference_matrix(ctx, oracle, xt, yt):
    x, y = rng_col(xt, 100, 1, lo=0, hi=9), rng_col(yt, 100, 2, lo=0, hi=9)
    with pytest.raises(ob.OracleError):
        oracle.cmp(ob.LT, xt, x, yt, y)
    with pytest.raises(capi.RfbError) as e:
        ctx.cmp(ob.LT, xt, dev(x), yt, dev(y))
    assert e.value.kind == "type"
    with pytest.raises(capi.RfbError) as e:
        ctx.cmp_where(ob.LT, xt, dev(x), 3) if xt in (ob.U8, ob.B8) else ctx.cmp(ob.LT, xt, dev(x), yt, 3)
    assert e.value.kind == "type"


@pytest.mark.parametrize("op", CMPS)
def test_cmp_date_vs_timestamp(ctx, oracle, op):
    # the date side is converted to nanoseconds (reference core/cmp.c:243-257, goldens tests/lang.c:3671-3678)
    n = 30_011
    r = np.random.default_rng(op)
    d = r.integers(8000, 8010, n).astype(np.int32)
    d[::19] = ob.NULL_I32
    ts = (r.integers(8000, 8010, n) * 86400_000_000_000 + r.integers(-1, 2, n) * 3600_000_000_000).astype(np.int64)
    ts[::23] = ob.NULL_I64
    assert np.array_equal(host(ctx.cmp(op, ob.DATE, dev(d), ob.TIMESTAMP, dev(ts))), oracle.cmp(op, ob.DATE, d, ob.TIMESTAMP, ts))
    assert np.array_equal(host(ctx.cmp(op, ob.TIMESTAMP, dev(ts), ob.DATE, dev(d))), oracle.cmp(op, ob.TIMESTAMP, ts, ob.DATE, d))
    assert np.array_equal(host(ctx.cmp(op, ob.DATE, dev(d), ob.TIMESTAMP, int(ts[1]))), oracle.cmp(op, ob.DATE, d, ob.TIMESTAMP, ts[1]))
    assert np.array_equal(host(ctx.cmp(op, ob.DATE, int(d[1]), ob.TIMESTAMP, dev(ts))), oracle.cmp(op, ob.DATE, d[1], ob.TIMESTAMP, ts))


def test_cmp_length_mismatch_and_unaligned(ctx, oracle):
    a, b = dev(np.arange(10, dtype=np.int64)), dev(np.arange(11, dtype=np.int64))
    with pytest.raises(capi.RfbError) as e:
        ctx.cmp(ob.LT, ob.I64, a, ob.I64, b)
    assert e.value.kind == "length"
    x = rng_col(ob.I64, 50_001, seed=1, lo=-5, hi=5)
    d = dev(x)
    assert np.array_equal(host(ctx.cmp(ob.GE, ob.I64, d[1:], ob.I64, 0)), oracle.cmp(ob.GE, ob.I64, x[1:], ob.I64, 0))


@pytest.mark.parametrize("n", [1, 17, 25001, 1_000_003])
def test_mask_logic_and_or_not(ctx, n):
    """and / or / not on masks (reference core/logic.c:34-86, core/order.c:422-443; golden shape tests/lang.c:2893-2897:
    a 25001-row `and` filter); any non-zero byte is true, results are 0/1"""
    r = np.random.default_rng(n)
    a = (r.random(n) < 0.5).astype(np.uint8) * r.integers(1, 255, n, dtype=np.uint8)
    b = (r.random(n) < 0.3).astype(np.uint8)
    da, db = dev(a), dev(b)
    assert np.array_equal(host(ctx.mask_logic(capi.M_AND, da, db)), ((a != 0) & (b != 0)).astype(np.uint8))
    assert np.array_equal(host(ctx.mask_logic(capi.M_OR, da, db)), ((a != 0) | (b != 0)).astype(np.uint8))
    assert np.array_equal(host(ctx.mask_logic(capi.M_NOT, da)), (a == 0).astype(np.uint8))
    assert np.array_equal(host(ctx.mask_logic(capi.M_AND, da, True)), (a != 0).astype(np.uint8))
    assert np.array_equal(host(ctx.mask_logic(capi.M_OR, da, False)), (a != 0).astype(np.uint8))
    assert not host(ctx.mask_logic(capi.M_AND, da, False)).any()
    if n > 20:
        with pytest.raises(capi.RfbError) as e:
            ctx.mask_logic(capi.M_AND, da, db[1:])
        assert e.value.kind == "length"


# ---------------------------------------------------------------- where / gather

@pytest.mark.parametrize("n", [1, 15, 16, 16383, 16384, 16385, 1_000_003])
@pytest.mark.parametrize("frac", [0.0, 0.01, 0.5, 1.0])
def test_where(ctx, oracle, n, frac):
    r = np.random.default_rng(n)
    mask = (r.random(n) < frac).astype(np.uint8)
    if frac == 0.5:
        mask[mask > 0] = r.integers(1, 255, int(mask.sum()), dtype=np.uint8)      # any non-zero byte selects
    assert np.array_equal(host(ctx.where(dev(mask))), oracle.where(mask))


@pytest.mark.parametrize("n", [8191, 16384, 16385, 3_000_017, 40_000_003])
def test_compaction_with_drifting_selectivity(ctx, oracle, n):
    """the compaction kernels at sizes around one tile and over thousands of tiles, with a selectivity that drifts between 0 and 1
    along the column (empty and full tiles, long look-back chains); the count and every id are checked"""
    r = np.random.default_rng(n % 1000)
    frac = r.random(n // 50_000 + 1).repeat(50_000)[:n]           # selectivity drifts between 0 and 1 along the column
    mask = (r.random(n) < frac).astype(np.uint8)
    mask[: min(n, 20_000)] = 0
    mask[-min(n, 9_000):] = 1
    want = np.flatnonzero(mask).astype(np.int64)
    assert np.array_equal(host(ctx.where(dev(mask))), want)
    x = np.where(mask != 0, 3, 5).astype(np.int64)
    x[::97] = ob.NULL_I64
    want = oracle.where(oracle.cmp(ob.LT, ob.I64, x, ob.I64, 4)) if n <= 3_000_017 else np.flatnonzero(x < 4).astype(np.int64)
    assert np.array_equal(host(ctx.cmp_where(ob.LT, ob.I64, dev(x), 4)), want)
    x32 = x.astype(np.int32)
    x32[::97] = ob.NULL_I32
    assert np.array_equal(host(ctx.cmp_where(ob.LT, ob.I32, dev(x32), 4)), want)
    del x32
    xf = np.where(mask != 0, 3.0, 5.0)
    xf[::97] = np.nan
    assert np.array_equal(host(ctx.cmp_where(ob.LT, ob.F64, dev(xf), 4.0)), want)


def test_where_unaligned_mask(ctx, oracle):
    mask = (np.random.default_rng(5).random(100_003) < 0.3).astype(np.uint8)
    assert np.array_equal(host(ctx.where(dev(mask)[5:])), oracle.where(mask[5:]))


@pytest.mark.parametrize("t", CMP_T + [ob.TIMESTAMP])
@pytest.mark.parametrize("op", [ob.LT, ob.EQ, ob.GE])
@pytest.mark.parametrize("n", [1, 4097, 300_001])
def test_cmp_where_equals_where_of_cmp(ctx, oracle, t, op, n):
    x = rng_col(t, n, seed=n + 3 * t, null_frac=0.03, lo=-30 if t != ob.U8 else 0, hi=30)
    if t == ob.F64:
        x = np.round(x)
    want = oracle.where(oracle.cmp(op, t, x, t, 4))
    assert np.array_equal(host(ctx.cmp_where(op, t, dev(x), 4)), want)


@pytest.mark.parametrize("t", ALL_T + [ob.TIMESTAMP, ob.DATE])
def test_gather(ctx, oracle, t):
    n = 123_457
    col = rng_col(t, n, seed=8, null_frac=0.02)
    ids = np.random.default_rng(9).integers(0, n, 77_777).astype(np.int64)
    got = host(ctx.gather(t, dev(col), dev(ids)))
    want = oracle.at_ids(t, col, ids)
    assert same_f64(got, want) if t == ob.F64 else np.array_equal(got, want)


# ---------------------------------------------------------------- element-wise arithmetic

def small_col(t, n, seed):
    a = rng_col(t, n, seed=seed, null_frac=0.04, lo=-50, hi=50)
    if t == ob.F64:
        a = np.round(a * 4) / 4
        a[::17] = 0.0
    else:
        a[::17] = 0          # division by zero -> null
    return a


def check_binop(got, got_t, want, want_t, op):
    assert got_t == want_t
    if want_t == ob.F64:
        assert same_f64(got, want, zero_sign=False, max_ulp=1 if op == ob.FDIV else 0)
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("op", ARITH)
@pytest.mark.parametrize("xt,yt", [(a, b) for a in (ob.I32, ob.I64, ob.F64) for b in (ob.I32, ob.I64, ob.F64)])
def test_binop_vector_vector(ctx, oracle, op, xt, yt):
    n = 50_003
    x, y = small_col(xt, n, 1), small_col(yt, n, 2)
    want, wt = oracle.binop(op, xt, x, yt, y)
    got, gt = ctx.binop(op, xt, dev(x), yt, dev(y))
    check_binop(host(got), gt, want, wt, op)


@pytest.mark.parametrize("op", ARITH)
@pytest.mark.parametrize("xt,yt", [(ob.I64, ob.I64), (ob.F64, ob.F64), (ob.I32, ob.I32), (ob.I64, ob.F64), (ob.F64, ob.I64),
                                   (ob.I32, ob.I64)])
@pytest.mark.parametrize("k", [3, -7, 0])
def test_binop_vector_atom_both_sides(ctx, oracle, op, xt, yt, k):
    n = 20_011
    x = small_col(xt, n, 3)
    kk = float(k) + (0.5 if yt == ob.F64 and k else 0.0) if yt == ob.F64 else k
    want, wt = oracle.binop(op, xt, x, yt, kk)
    got, gt = ctx.binop(op, xt, dev(x), yt, kk)
    check_binop(host(got), gt, want, wt, op)
    y = small_col(yt, n, 4)
    kx = float(k) if xt == ob.F64 else k
    want, wt = oracle.binop(op, xt, kx, yt, y)
    got, gt = ctx.binop(op, xt, kx, yt, dev(y))
    check_binop(host(got), gt, want, wt, op)


def test_binop_null_atoms_and_wraparound(ctx, oracle):
    x = np.array([1, ob.NULL_I64, ob.INF_I64, -5, ob.NULL_I64 + 1] * 100, np.int64)
    for op in ARITH:
        for k in (ob.NULL_I64, ob.INF_I64, -1, 2):
            want, wt = oracle.binop(op, ob.I64, x, ob.I64, k)
            got, gt = ctx.binop(op, ob.I64, dev(x), ob.I64, k)
            check_binop(host(got), gt, want, wt, op)
    x32 = np.array([1, ob.NULL_I32, 2 ** 31 - 1, -5, -(2 ** 31) + 1] * 100, np.int32)
    for op in ARITH:
        want, wt = oracle.binop(op, ob.I32, x32, ob.I32, x32[::-1].copy())
        got, gt = ctx.binop(op, ob.I32, dev(x32), ob.I32, dev(x32[::-1].copy()))
        check_binop(host(got), gt, want, wt, op)


@pytest.mark.parametrize("op", [ob.DIV, ob.MOD, ob.XBAR])
def test_division_by_a_constant_atom(ctx, oracle, op):
    """i64 column (/ | % | xbar) i64 atom runs on a host-computed magic multiplier instead of the 64-bit hardware division:
    floor semantics for / and %, truncation inside xbar, every sign combination, extreme operands and divisors"""
    r = np.random.default_rng(op)
    x = np.concatenate([r.integers(-(1 << 62), 1 << 62, 60_000), r.integers(-1000, 1000, 20_000),
                        np.array([0, 1, -1, ob.INF_I64, ob.NULL_I64 + 1, ob.NULL_I64, 1 << 62, -(1 << 62)])]).astype(np.int64)
    for k in (2, 3, -3, 7, -7, 17, 1000, 4096, -4096, 1_000_003, (1 << 31), (1 << 31) + 1, (1 << 62) + 5, ob.INF_I64, -ob.INF_I64, 1, -1, 0):
        want, wt = oracle.binop(op, ob.I64, x, ob.I64, k)
        got, gt = ctx.binop(op, ob.I64, dev(x), ob.I64, k)
        assert gt == wt and np.array_equal(host(got), want), k


@pytest.mark.parametrize("op", ARITH)
@pytest.mark.parametrize("xt", ALL_ARITH_TYPES)
def test_binop_full_type_matrix(ctx, oracle, op, xt):
    """every (operator, operand types, form) case of the reference outside I32/I64/F64 x I32/I64/F64 (core/math.c:251-1782: B8 /
    U8 / I16 / DATE / TIME / TIMESTAMP operands, unit conversions included): k_binop_typed against the pinned matrix restatement,
    a type error exactly where the reference has no case; lengths straddle the vector tile and leave a scalar tail"""
    plain = (ob.I32, ob.I64, ob.F64)
    for yt in ALL_ARITH_TYPES:
        if xt in plain and yt in plain:
            continue
        for n in (4_099, 70_001):
            x, y = typed_col(xt, n, 11 + n), typed_col(yt, n, 12 + n)
            dx, dy = dev(x), dev(y)
            for form, (a, b, da, db) in ((0, (x, y, dx, dy)), (1, (x, y[5], dx, y[5])), (1, (x, y[0], dx, y[0])), (1, (x, y[3], dx, y[3])),
                                         (2, (x[5], y, x[5], dy)), (2, (x[0], y, x[0], dy)), (2, (x[3], y, x[3], dy))):
                want_t = oracle.binop_form(op, form, xt, yt)
                assert ctx.lib.rfb_binop_type_form(op, form, xt, yt) == (want_t if want_t >= 0 else capi.ERR_TYPE)
                if want_t < 0:
                    with pytest.raises(capi.RfbError) as e:
                        ctx.binop(op, xt, da, yt, db)
                    assert e.value.kind == "type"
                    continue
                want, wt = oracle.binop(op, xt, a, yt, b)
                got, gt = ctx.binop(op, xt, da, yt, db)
                got = host(got)
                assert gt == wt == want_t, (op, form, xt, yt)
                if wt == ob.F64:
                    assert same_f64(got, want, zero_sign=False, max_ulp=1 if op == ob.FDIV else 0), (op, form, xt, yt)
                else:
                    assert np.array_equal(got, want), (op, form, xt, yt, np.flatnonzero(got != want)[:8])


@pytest.mark.parametrize("op", [ob.DIV, ob.MOD, ob.XBAR])
def test_timestamp_division_by_a_constant_atom(ctx, oracle, op):
    """TIMESTAMP (/ | % | xbar) by an I64 / TIMESTAMP atom shares the magic-multiplier kernel of i64 by an atom"""
    x = typed_col(ob.TIMESTAMP, 50_021, 3)
    for yt in (ob.I64, ob.TIMESTAMP):
        if oracle.binop_form(op, 1, ob.TIMESTAMP, yt) < 0:
            continue
        for k in (60_000_000_000, 86_400_000_000_000, -7, 3, 1, -1, 0, ob.NULL_I64):
            want, wt = oracle.binop(op, ob.TIMESTAMP, x, yt, k)
            got, gt = ctx.binop(op, ob.TIMESTAMP, dev(x), yt, k)
            assert gt == wt and np.array_equal(host(got), want), (yt, k)


@pytest.mark.parametrize("op", [ob.DIV, ob.MOD, ob.XBAR])
@pytest.mark.parametrize("xt", [ob.TIME, ob.DATE, ob.I32])
def test_32bit_division_by_a_constant_atom(ctx, oracle, op, xt):
    """TIME / DATE / I32 columns (/ | % | xbar) by an I32 / I64 atom: the 32-bit operator family on the magic-multiplier kernel
    (`xbar time 60000`); every sign combination, extreme operands, divisors up to 2^31 - 1, the atoms the generic kernel keeps"""
    r = np.random.default_rng(op * 10 + xt)
    x = np.concatenate([r.integers(-(1 << 31) + 1, (1 << 31) - 1, 60_000), r.integers(-1000, 1000, 20_000),
                        np.array([0, 1, -1, (1 << 31) - 1, -(1 << 31) + 1, ob.NULL_I32, 86_399_999, -86_400_000])]).astype(np.int32)
    for yt in (ob.I32, ob.I64):
        if oracle.binop_form(op, 1, xt, yt) < 0:
            continue
        for k in (60_000, 1000, 7, -7, 3, -3, 4096, -4096, 86_400_000, (1 << 31) - 1, -((1 << 31) - 1), 2, -2, 1, -1, 0, ob.NULL_I32 if yt == ob.I32 else ob.NULL_I64):
            want, wt = oracle.binop(op, xt, x, yt, k)
            got, gt = ctx.binop(op, xt, dev(x), yt, k)
            assert gt == wt and np.array_equal(host(got), want), (yt, k, np.flatnonzero(host(got) != want)[:5])
    if xt == ob.TIME:                # an I64 atom beyond 32 bits is narrowed like i64_to_time does it
        for k in ((1 << 32) + 60_000, -(1 << 40) + 17):
            if oracle.binop_form(op, 1, xt, ob.I64) >= 0:
                want, wt = oracle.binop(op, xt, x, ob.I64, k)
                got, gt = ctx.binop(op, xt, dev(x), ob.I64, k)
                assert gt == wt and np.array_equal(host(got), want), k


def test_binop_unaligned_typed_operands(ctx, oracle):
    """operand views that are not 16-byte aligned take the scalar path of k_binop_typed"""
    n = 30_001
    x, y = typed_col(ob.TIMESTAMP, n + 1, 5), typed_col(ob.TIME, n + 3, 6)
    dx, dy = dev(x)[1:], dev(y)[3:]
    for op in (ob.ADD, ob.SUB):
        want, wt = oracle.binop(op, ob.TIMESTAMP, x[1:], ob.TIME, y[3:])
        got, gt = ctx.binop(op, ob.TIMESTAMP, dx, ob.TIME, dy)
        assert gt == wt == ob.TIMESTAMP and np.array_equal(host(got), want)


def test_binop_errors(ctx):
    a, b = dev(np.zeros(4, np.int64)), dev(np.zeros(5, np.int64))
    with pytest.raises(capi.RfbError) as e:
        ctx.binop(ob.ADD, ob.I64, a, ob.I64, b)
    assert e.value.kind == "length"
    with pytest.raises(capi.RfbError) as e:
        ctx.binop(ob.MUL, ob.DATE, dev(np.zeros(4, np.int32)), ob.DATE, dev(np.zeros(4, np.int32)))   # no such case in the reference's matrix
    assert e.value.kind == "type"


def test_fused_expression_matches_operator_pipeline(ctx, oracle):
    """(avg (+ (* a b) c)): the fused kernel and the operator-at-a-time kernels agree bit for bit on the map part"""
    n = 100_003
    a, b, c = (rng_col(ob.F64, n, seed=s, null_frac=0.01, lo=0, hi=1) for s in (1, 2, 3))
    t1, _ = ctx.binop(ob.MUL, ob.F64, dev(a), ob.F64, dev(b))
    t2, _ = ctx.binop(ob.ADD, ob.F64, t1, ob.F64, dev(c))
    want = oracle.binop(ob.ADD, ob.F64, oracle.binop(ob.MUL, ob.F64, a, ob.F64, b)[0], ob.F64, c)[0]
    assert same_f64(host(t2), want)
    unfused = ctx.fold(capi.F_SUM | capi.F_CNT, ob.F64, t2, n)
    fused = ctx.fma_fold(capi.F_SUM | capi.F_CNT, dev(a), dev(b), dev(c), n)
    assert fused.nonnull == unfused.nonnull and fused.sum == unfused.sum


@pytest.mark.parametrize("op", [ob.ROUND, ob.FLOOR, ob.CEIL])
def test_unop_f64(ctx, oracle, op):
    x = np.concatenate([rng_col(ob.F64, 70_001, seed=op, null_frac=0.02, lo=-1e6, hi=1e6),
                        np.array([0.0, -0.0, 0.5, -0.5, 1.5, 2.5, -1.5, -2.5, 1e15 + 0.5, -1e15 - 0.5, np.nan, 4.0, -4.0])])
    assert same_f64(host(ctx.unop_f64(op, dev(x))), oracle.unop_f64(op, x), zero_sign=False)


# ---------------------------------------------------------------- group index + grouped aggregates

def keys_dense(n, card, seed):
    return (np.random.default_rng(seed).integers(0, card, n) * 3 - 1000).astype(np.int64) // 3 * 1  # shifted, contiguous-ish


@pytest.mark.parametrize("n,card", [(1, 1), (7, 3), (16385, 100), (300_007, 1000), (300_007, 250_000), (1_000_003, 100_000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_group_dense_matches_oracle_numbering(ctx, oracle, n, card, filtered):
    r = np.random.default_rng(n + card)
    keys = (r.integers(0, card, n) - 500).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    wg, wf, wi = oracle.group_i64(keys, filt)
    gg, gf, gi = ctx.group_i64(dev(keys), dev(filt) if filtered else None)
    assert (gi.groups, gi.dense, gi.index_type, gi.min, gi.max, gi.range) == (wi.groups, wi.dense, wi.index_type, wi.min, wi.max, wi.range)
    assert np.array_equal(host(gf), wf) and np.array_equal(host(gg), wg)


@pytest.mark.parametrize("n", [4, 1000, 200_003])
@pytest.mark.parametrize("filtered", [False, True])
def test_group_sparse_matches_oracle_numbering(ctx, oracle, n, filtered):
    r = np.random.default_rng(n)
    pool = r.integers(-(1 << 60), 1 << 60, max(2, n // 7)).astype(np.int64)
    keys = pool[r.integers(0, pool.shape[0], n)]
    filt = np.sort(r.choice(n, max(1, n // 2), replace=False)).astype(np.int64) if filtered else None
    wg, wf, wi = oracle.group_i64(keys, filt)
    gg, gf, gi = ctx.group_i64(dev(keys), dev(filt) if filtered else None)
    assert n < 100 or wi.dense == 0           # (a 2-row filter of 4 wide keys can still be "dense": range <= len)
    assert (gi.groups, gi.dense, gi.index_type) == (wi.groups, wi.dense, wi.index_type)
    assert np.array_equal(host(gf), wf) and np.array_equal(host(gg), wg)


@pytest.mark.parametrize("ncols", [2, 3, 6])
@pytest.mark.parametrize("filtered", [False, True])
@pytest.mark.parametrize("n", [5, 70_001, 500_003])
def test_group_multi_key_matches_oracle_numbering(ctx, oracle, ncols, filtered, n):
    """index_group_list (perfect-hash key fusion, reference core/index.c:2308-2424): first-occurrence numbering of key tuples"""
    r = np.random.default_rng(n + ncols)
    cols = [(r.integers(0, 3 + c, n) * (c + 1) - 7 * c).astype(np.int64) for c in range(ncols)]
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    wg, wf, groups = oracle.group_multi(cols, filt)
    gg, gf, gi = ctx.group_keys([dev(c) for c in cols], dev(filt) if filtered else None)
    assert gi.groups == groups
    assert np.array_equal(host(gf), wf) and np.array_equal(host(gg), wg)


@pytest.mark.parametrize("ncols", [2, 4])
@pytest.mark.parametrize("filtered", [False, True])
@pytest.mark.parametrize("n", [3, 70_001, 500_003])
def test_group_multi_key_row_hash_path(ctx, oracle, ncols, filtered, n):
    """index_group_list when the product of the key ranges does not fit an i64 (reference core/index.c:2556-2729: row
    hashes + open addressing): tuples are grouped through a table of representative rows; numbering = first occurrence
    (the reference's order at -c 1)"""
    r = np.random.default_rng(n * 3 + ncols)
    pools = [r.integers(-(1 << 61), 1 << 61, 5 + c).astype(np.int64) for c in range(ncols)]     # few distinct, huge ranges
    cols = [p[r.integers(0, p.shape[0], n)] for p in pools]
    cols[-1][::5] = ob.NULL_I64                                                                 # a null is a key like any other
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    wg, wf, groups = oracle.group_multi(cols, filt)
    gg, gf, gi = ctx.group_keys([dev(c) for c in cols], dev(filt) if filtered else None)
    assert gi.groups == groups and (gi.dense == 0 or n < 100)     # (a 1-row filter of 3 rows has key ranges of 1: perfect hash)
    assert np.array_equal(host(gf), wf) and np.array_equal(host(gg), wg)


def test_group_empty(ctx):
    gg, gf, gi = ctx.group_i64(dev(np.empty(0, np.int64)), None)
    assert gi.groups == 0 and gi.dense == 1 and gf.shape[0] == 0


AGGR_CASES = [(ob.SUM, ob.I64), (ob.SUM, ob.I16), (ob.SUM, ob.F64), (ob.MIN, ob.I64), (ob.MAX, ob.I64), (ob.MIN, ob.I16), (ob.MAX, ob.I16),
              (ob.MIN, ob.F64), (ob.MAX, ob.F64), (ob.MIN, ob.DATE), (ob.MAX, ob.TIME), (ob.MIN, ob.TIMESTAMP), (ob.COUNT, ob.I64),
              (ob.COUNT, ob.F64), (ob.COUNT, ob.I32), (ob.AVG, ob.I64), (ob.AVG, ob.I32), (ob.AVG, ob.I16), (ob.AVG, ob.F64), (ob.AVG, ob.TIME)]
A_OF = {ob.SUM: capi.A_SUM, ob.MIN: capi.A_MIN, ob.MAX: capi.A_MAX, ob.COUNT: capi.A_COUNT, ob.AVG: capi.A_AVG}


@pytest.mark.parametrize("op,vt", AGGR_CASES)
@pytest.mark.parametrize("filtered", [False, True])
@pytest.mark.parametrize("card", [777, 5000])      # CTA-private shared-memory accumulators / device-wide atomics
def test_aggr(ctx, oracle, op, vt, filtered, card):
    n = 200_003
    r = np.random.default_rng(op * 31 + vt)
    keys = r.integers(0, card, n).astype(np.int64)
    # I16 sums are accumulated in 16 bits by the reference and a running sum that lands exactly on 0x8000 turns the group
    # null from then on (sticky ADDI16, core/aggr.c:1083-1085) — an order-dependent artefact (it differs between the
    # reference's own thread counts); keep 16-bit partial sums far from wrapping so the result is order-free
    span = 60 if vt == ob.I16 else 1000
    val = rng_col(vt, n, seed=vt + op, null_frac=0.0005, lo=-span, hi=span)
    if vt == ob.F64:
        val = np.round(val * 8) / 8        # dyadic rationals: every partial sum is exact, any order gives the same bits
    filt = np.sort(r.choice(n, n // 2, replace=False)).astype(np.int64) if filtered else None
    wg, wf, wi = oracle.group_i64(keys, filt)
    want, wt = oracle.aggr(op, vt, val, wg, wi.groups, filt)
    got, gt = ctx.aggr(A_OF[op], vt, dev(val), dev(wg), wi.groups, dev(filt) if filtered else None)
    assert gt == wt
    assert same_f64(host(got), want, zero_sign=False) if wt == ob.F64 else np.array_equal(host(got), want)


@pytest.mark.parametrize("vt", [ob.I64, ob.F64, ob.I32, ob.I16, ob.TIME, ob.TIMESTAMP, ob.DATE, ob.U8])
@pytest.mark.parametrize("filtered", [False, True])
@pytest.mark.parametrize("n,card", [(9, 3), (16_385, 100), (200_003, 5000)])
def test_aggr_first_last(ctx, oracle, vt, filtered, n, card):
    """aggr_first (value at the group's first row, nulls included) and aggr_last (last non-null value of the first worker chunk
    that has one, for 1 / 3 / 16 chunks: the reference's answer at 1 / 3 / 16 executors) against the oracle, which is pinned
    against the compiled reference in tests/test_oracle_vs_reference.py::test_grouped_first_and_last"""
    r = np.random.default_rng(n + vt)
    keys = r.integers(0, card, n).astype(np.int64)
    val = rng_col(vt, n, seed=vt + 5, null_frac=0.4, lo=-1000 if vt != ob.U8 else 0, hi=1000 if vt != ob.U8 else 200)
    filt = np.sort(r.choice(n, max(1, n // 2), replace=False)).astype(np.int64) if filtered else None
    wg, wf, wi = oracle.group_i64(keys, filt)
    want, wt = oracle.aggr(ob.FIRST, vt, val, wg, wi.groups, filt)
    got, gt = ctx.aggr(capi.A_FIRST, vt, dev(val), dev(wg), wi.groups, dev(filt) if filtered else None)
    assert gt == wt and (same_f64(host(got), want, zero_sign=False) if wt == ob.F64 else np.array_equal(host(got), want))
    if vt == ob.U8:
        with pytest.raises(capi.RfbError):
            ctx.aggr(capi.A_LAST, vt, dev(val), dev(wg), wi.groups)
        return
    for chunks in (1, 3, 16):
        want, wt = oracle.aggr_last(vt, val, wg, wi.groups, chunks, filt)
        got, gt = ctx.aggr_last(vt, dev(val), dev(wg), wi.groups, chunks, dev(filt) if filtered else None)
        assert gt == wt and (same_f64(host(got), want, zero_sign=False) if wt == ob.F64 else np.array_equal(host(got), want)), chunks
    want, _ = oracle.aggr(ob.LAST, vt, val, wg, wi.groups, filt)
    got, _ = ctx.aggr(capi.A_LAST, vt, dev(val), dev(wg), wi.groups, dev(filt) if filtered else None)
    assert same_f64(host(got), want, zero_sign=False) if wt == ob.F64 else np.array_equal(host(got), want)


@pytest.mark.parametrize("op", [ob.SUM, ob.AVG])
@pytest.mark.parametrize("card,vbits,nulls", [(100_000, 20, 0.0005), (250_000, 19, 0.0), (60_000, 45, 0.0005), (20_000, 62, 0.0), (100_000, 20, 0.2)])
def test_aggr_through_partition_passes(ctx, oracle, monkeypatch, op, card, vbits, nulls):
    """aggr_sum / aggr_avg of i64 values at 1e4 .. 2.6e5 groups take the narrow partition passes of the fused group-by (group id
    = key; packed 32 / 64-bit records, nulls and wide values through the exception list); values that fit no record format
    (62 bits) and null-heavy columns fall back to the device-wide atomics: same bits either way"""
    if op == ob.AVG and vbits > 45:
        pytest.skip("the oracle averages in f64 like the reference: only sums below 2^53 are order-free")
    monkeypatch.setenv("RFB_PART_MIN_ROWS", "1000")
    n = 1_200_007
    r = np.random.default_rng(card + vbits)
    gid = r.integers(0, card, n).astype(np.int64)
    gid[:card] = np.arange(card)                     # every group occurs
    val = r.integers(-(1 << (vbits - 1)), 1 << (vbits - 1), n).astype(np.int64) if vbits > 32 else r.integers(0, 1 << vbits, n).astype(np.int64)
    val[r.random(n) < nulls] = ob.NULL_I64
    want, wt = oracle.aggr(op, ob.I64, val, gid, card)
    launches = ctx.launches
    got, gt = ctx.aggr(A_OF[op], ob.I64, dev(val), dev(gid), card)
    assert gt == wt
    assert same_f64(host(got), want, zero_sign=False) if wt == ob.F64 else np.array_equal(host(got), want)
    if vbits <= 45 and nulls < 0.01:
        assert ctx.launches - launches >= 8          # census x3, scatter, (tail), accumulate, exceptions, finalise: the partition path ran


def test_aggr_sticky_null_and_all_null_group(ctx, oracle):
    # grouped sum over [1 0Nl | 3 4] -> [0Nl 7]; count -> [2 2]; avg -> [1.0 3.5]; min/max -> [1 3]/[1 4]  (SURVEY §8a probes)
    gid = np.array([0, 0, 1, 1, 2], np.int64)
    val = np.array([1, ob.NULL_I64, 3, 4, ob.NULL_I64], np.int64)
    for op in (ob.SUM, ob.COUNT, ob.AVG, ob.MIN, ob.MAX):
        want, wt = oracle.aggr(op, ob.I64, val, gid, 3)
        got, gt = ctx.aggr(A_OF[op], ob.I64, dev(val), dev(gid), 3)
        assert gt == wt
        assert same_f64(host(got), want) if wt == ob.F64 else np.array_equal(host(got), want), op
    s, _ = ctx.aggr(capi.A_SUM, ob.I64, dev(val), dev(gid), 3)
    assert host(s).tolist() == [ob.NULL_I64, 7, ob.NULL_I64]


@pytest.mark.parametrize("op,vt", [(ob.SUM, ob.I32), (ob.SUM, ob.TIME), (ob.SUM, ob.TIMESTAMP), (ob.MIN, ob.I32), (ob.MAX, ob.I32),
                                   (ob.AVG, ob.TIMESTAMP), (ob.COUNT, ob.I16), (ob.COUNT, ob.U8)])
def test_aggr_type_errors_match_the_reference_drivers(ctx, oracle, op, vt):
    # the reference's non-parted aggr_* switch tables (core/aggr.c:1107-1150, 1152-1315, 2013-2133); pinned against the
    # compiled reference in tests/test_oracle_vs_reference.py::test_grouped_aggregate_type_errors
    gid, val = np.zeros(4, np.int64), rng_col(vt, 4, 1, lo=0, hi=9)
    with pytest.raises(ob.OracleError):
        oracle.aggr(op, vt, val, gid, 1)
    with pytest.raises(capi.RfbError) as e:
        ctx.aggr(A_OF[op], vt, dev(val), dev(gid), 1)
    assert e.value.kind == "type"


@pytest.mark.parametrize("key_type", [ob.I64, ob.I32])
@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("n,card", [(1, 1), (100_003, 50), (1_000_003, 100_000)])
def test_fused_group_sum_count(ctx, oracle, key_type, with_pred, n, card):
    r = np.random.default_rng(n + card)
    keys64 = (r.integers(0, card, n) + 17).astype(np.int64)
    val = r.integers(0, 1 << 20, n).astype(np.int64)
    val[r.random(n) < 0.0002] = ob.NULL_I64
    keys = keys64.astype(ob.NP_OF[key_type])
    if with_pred:
        k = 1 << 19
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, k))
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5, cmp_op=capi.LT, pred_type=ob.I64, pred=dev(val), k=k)
    else:
        filt = None
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5)
    wg, wf, wi = oracle.group_i64(keys64, filt)          # oracle path: I64 keys (the reference has no I32 grouping)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(host(gk), keys64[rows])         # group key column = key at each group's first row, first-occurrence order
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])


def test_fused_group_late_first_occurrence(ctx, oracle):
    """device-wide path: first rows are claimed on a growing row prefix; a key that only shows up in the last rows must
    still get its (late) first row and its place in the first-occurrence order"""
    n, card = 1_000_003, 2000
    r = np.random.default_rng(5)
    keys = r.integers(0, card - 3, n).astype(np.int64)
    keys[n - 1] = card - 1
    keys[n // 2] = card - 2
    keys[70_000] = card - 3
    val = r.integers(0, 1000, n).astype(np.int64)
    gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), card)
    wg, wf, wi = oracle.group_i64(keys)
    assert np.array_equal(host(gk), keys[wf]) and host(gk)[-1] == card - 1
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups)[0])


@pytest.mark.parametrize("strategy", ["smem", "part", "l2", "auto"])
@pytest.mark.parametrize("key_type", [ob.I64, ob.I32])
@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("n,card,kmin,skew", [
    (300_007, 3000, -1500, False),         # negative keys; one partition boundary inside the range
    (1_000_003, 100_000, 17, False),       # config-4 shape: 13 partitions of 8192 keys
    (1_000_003, 100_000, -8192 * 3 - 5, True),   # skewed (most rows in two keys), partitions of very different sizes
    (2_500_001, 1_900_000, 5, False),      # ~232 partitions, many slots never hit
])
def test_fused_group_strategies(ctx, oracle, monkeypatch, strategy, key_type, with_pred, n, card, kmin, skew):
    """the three accumulate strategies of the fused group-by (CTA-private shared memory / key-range partitions / L2 atomics)
    must agree with the oracle bit for bit: wrapping sums of negative and large values (32-bit carry chains in shared
    memory), sticky nulls, counts, first-occurrence order"""
    if strategy == "smem" and card > 8192:
        pytest.skip("shared-memory strategy needs range <= 8192")
    if strategy == "auto":      # what a large column gets: strategy picked from a row sample (scope-free residue / partition passes)
        monkeypatch.setenv("RFB_PART_MIN_ROWS", "1000")
    else:
        monkeypatch.setenv("RFB_GROUP_STRATEGY", strategy)
    r = np.random.default_rng(n + card + (1 if skew else 0))
    keys64 = r.integers(0, card, n).astype(np.int64)
    if skew:
        hot = r.random(n) < 0.8
        keys64[hot] = np.where(r.random(int(hot.sum())) < 0.5, 3, card - 2)
    keys64 += kmin
    val = r.integers(-(1 << 40), 1 << 40, n).astype(np.int64)
    val[::7] = r.integers(0, 1 << 20, val[::7].shape[0])
    val[::1013] = np.iinfo(np.int64).max       # sums wrap mod 2^64
    val[r.random(n) < 0.0002] = ob.NULL_I64
    keys = keys64.astype(ob.NP_OF[key_type])
    if with_pred:
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, 1 << 19))
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5, cmp_op=capi.LT, pred_type=ob.I64, pred=dev(val), k=1 << 19)
    else:
        filt = None
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5)
    wg, wf, wi = oracle.group_i64(keys64, filt)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(host(gk), keys64[rows])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])


@pytest.mark.parametrize("outlier", [5_000_000, 9000])
def test_fused_group_sample_misjudges_the_key_range(ctx, oracle, monkeypatch, outlier):
    """the partitioned strategy is chosen from a row sample (head / middle / tail windows).  Keys outside the windows can
    widen the range beyond what it can hold (-> the scatter is discarded, device-wide atomics take over) or beyond what
    the sample promised (-> more partitions than expected): the result must not depend on the guess"""
    monkeypatch.setenv("RFB_PART_MIN_ROWS", "1000")
    n = 600_011
    r = np.random.default_rng(outlier)
    keys = r.integers(0, 20_000 if outlier > 100_000 else 5000, n).astype(np.int64)
    keys[100_000] = outlier
    keys[400_000] = -outlier
    val = r.integers(-5000, 5000, n).astype(np.int64)
    gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), 30_000)
    wg, wf, wi = oracle.group_i64(keys)
    assert np.array_equal(host(gk), keys[wf])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups)[0])


@pytest.mark.parametrize("n,card,kmin,skew", [(1_000_003, 100_000, 17, False), (1_000_003, 100_000, -8192 * 3 - 5, True), (2_500_001, 1_900_000, 5, False),
                                              (300_007, 9000, -1500, False)])
@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("tma", ["1", "0"])
def test_fused_group_partitioned_tma_accumulate(ctx, oracle, monkeypatch, n, card, kmin, skew, with_pred, tma):
    """both variants of the accumulate pass — partition data staged by the TMA unit (cp.async.bulk + mbarrier ring, the default)
    and by 128-bit register loads (RFB_ACCUM_TMA=0) — against the oracle"""
    monkeypatch.setenv("RFB_GROUP_STRATEGY", "part")
    monkeypatch.setenv("RFB_ACCUM_TMA", tma)
    r = np.random.default_rng(n + card)
    keys = r.integers(0, card, n).astype(np.int64)
    if skew:
        hot = r.random(n) < 0.8
        keys[hot] = np.where(r.random(int(hot.sum())) < 0.5, 3, card - 2)
    keys += kmin
    val = r.integers(-(1 << 40), 1 << 40, n).astype(np.int64)
    val[r.random(n) < 0.0002] = ob.NULL_I64
    if with_pred:
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, 1 << 19))
        gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), card + 5, cmp_op=capi.LT, pred_type=ob.I64, pred=dev(val), k=1 << 19)
    else:
        filt = None
        gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), card + 5)
    wg, wf, wi = oracle.group_i64(keys, filt)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(host(gk), keys[rows])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])


@pytest.mark.parametrize("mode", ["narrow", "auto"])
@pytest.mark.parametrize("key_type", [ob.I64, ob.I32])
@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("n,card,kmin,vbits,skew", [
    (1_000_003, 100_000, 17, 20, False),          # config-4 shape: 25 partitions of 4096 keys, 32-bit records (12-bit slot | 20-bit value)
    (1_000_003, 100_000, -4096 * 7 - 5, 20, True),  # negative keys, skewed partitions
    (1_500_001, 250_000, 3, 19, False),           # 31 partitions of 8192 keys, 32-bit records (13 | 19)
    (1_000_003, 200_000, -70_000, 45, False),     # values need the 64-bit record (16-bit slot field | 48-bit signed value)
    (300_007, 9000, 5, 20, False),                # a range the one-pass residue strategy would normally take
])
def test_fused_group_narrow_records(ctx, oracle, monkeypatch, mode, key_type, with_pred, n, card, kmin, vbits, skew):
    """<= 32-partition path of the fused group-by: ballot ranking, packed 32/64-bit records, TMA-staged accumulate, and the
    exception side list (nulls, values that do not fit the record) against the oracle, bit for bit"""
    if mode == "narrow":
        monkeypatch.setenv("RFB_GROUP_STRATEGY", "narrow")
    else:
        monkeypatch.setenv("RFB_PART_MIN_ROWS", "1000")
    r = np.random.default_rng(n + card + vbits)
    keys64 = r.integers(0, card, n).astype(np.int64)
    if skew:
        hot = r.random(n) < 0.8
        keys64[hot] = np.where(r.random(int(hot.sum())) < 0.5, 3, card - 2)
    keys64 += kmin
    if vbits > 32:
        val = r.integers(-(1 << (vbits - 1)), 1 << (vbits - 1), n).astype(np.int64)
    else:
        val = r.integers(0, 1 << vbits, n).astype(np.int64)
    exc = r.random(n) < 0.0005                      # exceptions: nulls, negative and huge values
    val[exc] = r.choice(np.array([ob.NULL_I64, -1, -(1 << 50), (1 << 62) + 12345, np.iinfo(np.int64).max], dtype=np.int64), int(exc.sum()))
    keys = keys64.astype(ob.NP_OF[key_type])
    if with_pred:
        kk = 1 << (vbits - 1)
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, kk))
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5, cmp_op=capi.LT, pred_type=ob.I64, pred=dev(val), k=kk)
    else:
        filt = None
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5)
    wg, wf, wi = oracle.group_i64(keys64, filt)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(host(gk), keys64[rows])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])


@pytest.mark.parametrize("case", ["values", "keys"])
def test_fused_group_narrow_sample_misjudges(ctx, oracle, monkeypatch, case):
    """the record format and the 32-partition span are guessed from three row windows.  Values that only turn wide outside the
    windows overflow the exception list (the pass aborts, the 256-partition pass takes over); keys outside the windows can
    span more than 32 partitions (the scatter is void, the exact bounds it found are kept): the result must not change"""
    monkeypatch.setenv("RFB_PART_MIN_ROWS", "1000")
    n, card = 1_200_007, 60_000
    r = np.random.default_rng(77)
    keys = r.integers(0, card, n).astype(np.int64)
    val = r.integers(0, 1 << 18, n).astype(np.int64)
    if case == "values":
        val[100_000:500_000] = r.integers(-(1 << 60), 1 << 60, 400_000)
    else:
        keys[100_000:101_000] += 500_000
    gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), card + 1000)
    wg, wf, wi = oracle.group_i64(keys)
    assert np.array_equal(host(gk), keys[wf])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups)[0])


def _sparse_keys(r, n, card, key_type):
    """`card` distinct keys scattered over the whole key width (no dense domain), drawn uniformly"""
    if key_type == ob.I32:
        pool = r.permutation(np.unique(r.integers(-(1 << 31) + 1, 1 << 31, 2 * card).astype(np.int64)))[:card]
    else:
        pool = np.unique(r.integers(-(1 << 62), 1 << 62, card).astype(np.int64))
    return pool[r.integers(0, pool.shape[0], n)]


@pytest.mark.parametrize("key_type", [ob.I64, ob.I32])
@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("n,card", [(1, 1), (5000, 7), (700_001, 3000), (2_000_003, 40_000), (1_500_001, 400_000)])
def test_fused_group_sparse_keys(ctx, oracle, key_type, with_pred, n, card):
    """sparse key domains (range > 2^28): per-CTA open-addressing tables in shared memory fed by TMA-staged tiles, spills into
    the device-wide table (the 40k / 400k-key cases overflow the 8192-slot tables many times), groups in first-occurrence
    order (= the reference at -c 1; at -c N its order is hash-slot order): sums, sticky nulls, counts bit-exact"""
    r = np.random.default_rng(n + card)
    card = min(card, n)
    keys64 = _sparse_keys(r, n, card, key_type)
    val = r.integers(-(1 << 40), 1 << 40, n).astype(np.int64)
    val[::1013] = np.iinfo(np.int64).max
    val[r.random(n) < 0.0003] = ob.NULL_I64
    keys = keys64.astype(ob.NP_OF[key_type])
    if with_pred:
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, 1 << 19))
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5, cmp_op=capi.LT, pred_type=ob.I64, pred=dev(val), k=1 << 19)
    else:
        filt = None
        gk, gs, gc = ctx.group_sum_count(key_type, dev(keys), dev(val), card + 5)
    wg, wf, wi = oracle.group_i64(keys64, filt)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(host(gk), keys64[rows])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])


def test_fused_group_sparse_null_key(ctx):
    """NULL_I64 as a KEY: the reference's open-addressing table uses it as the empty marker and cannot hold it (core/hash.c:35-56,
    SURVEY a13); the device gives it a dedicated slot, so it groups like any other key.  Checked against numpy."""
    n = 300_007
    r = np.random.default_rng(9)
    pool = np.unique(r.integers(-(1 << 62), 1 << 62, 50).astype(np.int64))
    keys = pool[r.integers(0, pool.shape[0], n)]
    keys[r.integers(1000, n, 40)] = ob.NULL_I64
    val = r.integers(-1000, 1000, n).astype(np.int64)
    gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), 100)
    uk, first, inv, cnt = np.unique(keys, return_index=True, return_inverse=True, return_counts=True)
    order = np.argsort(first, kind="stable")                  # first-occurrence order
    sums = np.zeros(uk.shape[0], np.int64)
    np.add.at(sums, inv, val)
    assert np.array_equal(host(gk), uk[order]) and host(gk)[-1] == ob.NULL_I64
    assert np.array_equal(host(gs), sums[order]) and np.array_equal(host(gc), cnt[order])


def test_fused_group_hash_path_on_dense_keys_unaligned_and_overflow(ctx, oracle, monkeypatch):
    """the hash path forced onto a dense domain must agree with the dense strategies; columns that start at an odd element
    take the plain-load tile loader instead of the TMA ring; more groups than the output has room for is an error"""
    monkeypatch.setenv("RFB_GROUP_STRATEGY", "hash")
    n, card = 600_011, 20_000
    r = np.random.default_rng(3)
    keys = r.integers(0, card, n + 1).astype(np.int64)
    val = r.integers(-1000, 1000, n + 1).astype(np.int64)
    dk, dv = dev(keys), dev(val)
    for off in (0, 1):
        gk, gs, gc = ctx.group_sum_count(ob.I64, dk[off:], dv[off:], card)
        wg, wf, wi = oracle.group_i64(keys[off:])
        assert np.array_equal(host(gk), keys[off:][wf])
        assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val[off:], wg, wi.groups)[0])
        assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val[off:], wg, wi.groups)[0])
    with pytest.raises(capi.RfbError):
        ctx.group_sum_count(ob.I64, dk, dv, 1000)


def test_fused_group_partitioned_unaligned_columns(ctx, oracle, monkeypatch):
    """columns that start at an odd element (no 16-byte alignment) take the scalar tile loader"""
    monkeypatch.setenv("RFB_GROUP_STRATEGY", "part")
    n, card = 400_001, 20_000
    r = np.random.default_rng(11)
    keys = r.integers(0, card, n + 1).astype(np.int32)
    val = r.integers(-1000, 1000, n + 1).astype(np.int64)
    dk, dv = dev(keys), dev(val)
    gk, gs, gc = ctx.group_sum_count(ob.I32, dk[1:], dv[1:], card)
    k64 = keys[1:].astype(np.int64)
    wg, wf, wi = oracle.group_i64(k64)
    assert np.array_equal(host(gk), k64[wf])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val[1:], wg, wi.groups)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val[1:], wg, wi.groups)[0])


@pytest.mark.parametrize("card", [1, 2, 3])
@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("auto", [False, True])
def test_fused_group_single_and_few_keys(ctx, oracle, monkeypatch, card, with_pred, auto):
    """one to three keys: whole warps hit the same accumulator slot (the warp-aggregated path of the shared-memory strategies);
    wrapping sums, nulls in some warps only, long runs of one key"""
    if auto:
        monkeypatch.setenv("RFB_PART_MIN_ROWS", "1000")
    n = 400_003
    r = np.random.default_rng(card)
    keys = (np.arange(n) // 5000 % card + 7).astype(np.int64)          # runs of 5000 equal keys: most warps are uniform
    keys[r.random(n) < 0.001] = 7 + card - 1
    val = r.integers(-(1 << 62), 1 << 62, n).astype(np.int64)
    val[r.random(n) < 0.00005] = ob.NULL_I64
    if with_pred:
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, 1 << 61))
        gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), 8, cmp_op=capi.LT, pred_type=ob.I64, pred=dev(val), k=1 << 61)
    else:
        filt = None
        gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val), 8)
    wg, wf, wi = oracle.group_i64(keys, filt)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(host(gk), keys[rows])
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(host(gc), oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])
    # no nulls at all: the sums themselves (not the sticky null) are compared
    val2 = np.where(val == ob.NULL_I64, 5, val)
    gk, gs, gc = ctx.group_sum_count(ob.I64, dev(keys), dev(val2), 8)
    wg, wf, wi = oracle.group_i64(keys)
    assert np.array_equal(host(gs), oracle.aggr(ob.SUM, ob.I64, val2, wg, wi.groups)[0])


def test_fused_group_nothing_selected(ctx):
    keys, val = dev(np.arange(100, dtype=np.int64)), dev(np.arange(100, dtype=np.int64))
    gk, gs, gc = ctx.group_sum_count(ob.I64, keys, val, 10, cmp_op=capi.LT, pred_type=ob.I64, pred=val, k=-5)
    assert gk.shape[0] == 0


def test_group_merge_over_peer_memory_single_rank(ctx):
    """rfb_group_merge_peers with a world of one (the exchange buffer bound to itself): the publish / meet / fold / emit kernels on a
    list WITH repeated keys, so the merge has real work — first-occurrence order, wrapping sums, the sticky null of ADDI64, counts;
    three rounds reuse both halves of the buffer; a wide key domain is declined; an oversized list fails instead of hanging.
    (The multi-rank form is tests/test_gpu_nccl.py.)"""
    import ctypes as C
    h = (C.c_char * 64)()
    capi.check(ctx.lib.rfb_peer_groups_create(ctx.h, 50_000, h))
    capi.check(ctx.lib.rfb_peer_groups_bind(ctx.h, 0, 1, C.c_char_p(bytes(h.raw))))
    r = np.random.default_rng(4)
    for rnd in range(3):
        n = 40_000 - rnd
        keys = r.integers(-700, 9_000, n).astype(np.int64)
        sums = r.integers(-(1 << 62), 1 << 62, n).astype(np.int64)
        sums[::13] = ob.NULL_I64
        counts = r.integers(0, 1000, n).astype(np.int64)
        gk, gs, gc = ctx.group_merge_peers(dev(keys), dev(sums), dev(counts), 20_000)
        uk, first, inv = np.unique(keys, return_index=True, return_inverse=True)
        order = np.argsort(first, kind="stable")
        want_k = uk[order]
        tot = np.zeros(uk.shape[0], np.uint64)
        np.add.at(tot, inv, np.where(sums == ob.NULL_I64, 0, sums).astype(np.uint64))
        nul = np.zeros(uk.shape[0], bool)
        np.logical_or.at(nul, inv, sums == ob.NULL_I64)
        want_s = np.where(nul, np.int64(ob.NULL_I64), tot.astype(np.int64))[order]
        cnt = np.zeros(uk.shape[0], np.int64)
        np.add.at(cnt, inv, counts)
        assert np.array_equal(host(gk), want_k) and np.array_equal(host(gs), want_s) and np.array_equal(host(gc), cnt[order])
    wide = dev(np.array([0, 1 << 40], np.int64))
    with pytest.raises(capi.RfbError) as e:
        ctx.group_merge_peers(wide, wide, wide, 16)
    assert e.value.kind == "type"
    big = dev(np.arange(50_001, dtype=np.int64))
    with pytest.raises(capi.RfbError) as e:
        ctx.group_merge_peers(big, big, big, 60_000)
    assert e.value.kind == "arg"
    gk, gs, gc = ctx.group_merge_peers(dev(np.array([5, 3, 5], np.int64)), dev(np.array([1, 2, 3], np.int64)), dev(np.array([1, 1, 1], np.int64)), 16)
    assert host(gk).tolist() == [5, 3] and host(gs).tolist() == [4, 2] and host(gc).tolist() == [2, 1]      # still usable after the failures


# ---------------------------------------------------------------- key sort

@pytest.mark.parametrize("t", ALL_T + [ob.TIMESTAMP])
@pytest.mark.parametrize("desc", [False, True])
@pytest.mark.parametrize("n", [1, 2, 33, 2048, 2049, 100_003])
def test_sort_permutation(ctx, oracle, t, desc, n):
    col = rng_col(t, n, seed=n + t, null_frac=0.05, lo=-40 if t != ob.U8 else 0, hi=40)     # many duplicates: stability matters
    if t == ob.F64:
        col = np.round(col)
        col[::11] = -0.0
    assert np.array_equal(host(ctx.sort(t, dev(col), desc)), oracle.sort(t, col, desc))


@pytest.mark.parametrize("desc", [False, True])
def test_sort_wide_keys(ctx, oracle, desc):
    r = np.random.default_rng(3)
    col = r.integers(-(1 << 62), 1 << 62, 300_007).astype(np.int64)
    col[::101] = ob.NULL_I64
    assert np.array_equal(host(ctx.sort(ob.I64, dev(col), desc)), oracle.sort(ob.I64, col, desc))
    f = r.standard_normal(200_003) * 1e10
    f[::97] = np.nan
    f[1::97] = np.inf
    f[2::97] = -np.inf
    assert np.array_equal(host(ctx.sort(ob.F64, dev(f), desc)), oracle.sort(ob.F64, f, desc))


@pytest.mark.parametrize("algo", ["lsd", "onesweep"])
@pytest.mark.parametrize("desc", [False, True])
def test_sort_pass_structures_agree(ctx, oracle, monkeypatch, algo, desc):
    """both pass structures (single-sweep passes with decoupled look-back: the default; histogram + scatter passes: what columns of
    2^32 rows and more take) against the oracle, at sizes that span many tiles, with heavy duplicates (stability) and with digit
    columns that are constant (skipped passes)"""
    monkeypatch.setenv("RFB_SORT_ALGO", algo)
    r = np.random.default_rng(17)
    n = 1_000_003
    for col, t in ((r.integers(-5, 5, n).astype(np.int64), ob.I64), (r.integers(-(1 << 62), 1 << 62, n).astype(np.int64), ob.I64),
                   (r.integers(0, 1 << 20, n).astype(np.int64) << 24, ob.I64), (r.integers(-30000, 30000, n).astype(np.int32), ob.I32),
                   (np.round(r.standard_normal(n) * 100) / 4, ob.F64), (r.integers(0, 256, n).astype(np.uint8), ob.U8),
                   (r.integers(-300, 300, n).astype(np.int16), ob.I16)):
        assert np.array_equal(host(ctx.sort(t, dev(col), desc)), oracle.sort(t, col, desc)), (algo, t)


@pytest.mark.parametrize("n", [4095, 4096, 4097, 6143, 6144, 6145, 12288, 24577, 49_153])
def test_sort_around_tile_boundaries(ctx, oracle, n):
    """lengths around the tiles of the single-sweep passes (4096 rows for 64-bit key words, 6144 for 32-bit ones): the last tile is
    the only one that is not staged by the TMA unit"""
    r = np.random.default_rng(n)
    for col, t in ((r.integers(-(1 << 62), 1 << 62, n).astype(np.int64), ob.I64), (r.integers(-(1 << 30), 1 << 30, n).astype(np.int32), ob.I32),
                   (r.integers(0, 1 << 20, n).astype(np.int64), ob.I64), (r.standard_normal(n), ob.F64), (r.integers(-300, 300, n).astype(np.int16), ob.I16)):
        assert np.array_equal(host(ctx.sort(t, dev(col))), oracle.sort(t, col, False)), t
        assert np.array_equal(host(ctx.sort(t, dev(col), True)), oracle.sort(t, col, True)), t


@pytest.mark.parametrize("desc", [False, True])
def test_sort_narrow_windows_of_wide_columns(ctx, oracle, desc):
    """8-byte columns whose varying bytes fit a 32-bit window travel as 32-bit key words (window at byte 0, in the middle, at the
    top; exactly 4 varying bytes; 5 varying bytes take the 64-bit words again); doubles of a narrow range do the same"""
    r = np.random.default_rng(23)
    n = 300_017
    cols = [(r.integers(0, 60_000, n).astype(np.int64), ob.I64),
            (r.integers(0, 1 << 32, n).astype(np.int64), ob.I64),
            (r.integers(0, 1 << 33, n).astype(np.int64), ob.I64),
            ((r.integers(0, 1 << 30, n).astype(np.int64) << 16) + 77, ob.I64),
            (np.int64(1_700_000_000_000_000_000) + r.integers(0, 3_000_000_000, n).astype(np.int64), ob.TIMESTAMP),
            (-(r.integers(1, 1 << 28, n).astype(np.int64) << 32), ob.I64),
            (r.integers(0, 1000, n).astype(np.float64), ob.F64),
            (1e6 + r.integers(0, 4000, n) / 8.0, ob.F64)]
    for col, t in cols:
        col[::1001] = col[0]                               # duplicates: stability
        assert np.array_equal(host(ctx.sort(t, dev(col), desc)), oracle.sort(t, col, desc)), t


def test_sort_reference_goldens(ctx):
    # SURVEY §8a probes of the reference: NaN first, -0.0 before +0.0; both directions stable
    f = np.array([1.0, np.nan, -0.0, 0.0, -1.0])
    assert host(ctx.sort(ob.F64, dev(f))).tolist() == [1, 4, 2, 3, 0]
    k = np.array([2, 1, 2, 1, 2], np.int64)
    assert host(ctx.sort(ob.I64, dev(k))).tolist() == [1, 3, 0, 2, 4]
    assert host(ctx.sort(ob.I64, dev(k), True)).tolist() == [0, 2, 4, 1, 3]


def test_sort_constant_and_sorted_input(ctx, oracle):
    c = np.full(10_000, 7, np.int64)
    assert np.array_equal(host(ctx.sort(ob.I64, dev(c))), np.arange(10_000))
    s = np.arange(50_000, dtype=np.int64)
    assert np.array_equal(host(ctx.sort(ob.I64, dev(s), True)), s[::-1])
