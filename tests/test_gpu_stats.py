"""GPU parity for the order-statistic / list-valued aggregates (k_stats.cu): ray_med, ray_dev, aggr_med, aggr_dev,
aggr_row / aggr_collect against the CPU oracle (itself pinned against the compiled reference in
tests/test_oracle_vs_reference.py).  Medians, row lists and collected values are bit-exact; deviations are sums of squares in
fp64 whose order the reference leaves to its thread count: exact on inputs with exactly representable partial sums, else
held to 1e-12 relative."""
import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200 import capi
from tests.util import dev, host, rng_col, same_f64

pytestmark = pytest.mark.gpu


def close(a, b, rel=1e-12):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.allclose(a[~np.isnan(a)], b[~np.isnan(b)], rtol=rel, atol=0)


@pytest.mark.parametrize("t", [ob.U8, ob.I16, ob.I64])
@pytest.mark.parametrize("n", [1, 2, 5, 4097, 300_001])
@pytest.mark.parametrize("nulls", [False, True])
def test_med(ctx, oracle, t, n, nulls):
    col = rng_col(t, n, n + t, null_frac=0.2 if nulls else 0.0, lo=-50 if t != ob.U8 else 0, hi=50)
    assert same_f64([ctx.med(t, dev(col))], [oracle.med(t, col)])


def test_med_all_null_empty_and_type_errors(ctx, oracle):
    assert np.isnan(ctx.med(ob.I64, dev(np.full(9, ob.NULL_I64, np.int64))))
    assert np.isnan(ctx.med(ob.I64, dev(np.empty(0, np.int64))))
    for t in (ob.I32, ob.F64, ob.DATE):       # ray_med has no vector case for these (core/math.c:2555-2590)
        with pytest.raises(ob.OracleError):
            oracle.med(t, rng_col(t, 5, 1, lo=0, hi=9))
        with pytest.raises(capi.RfbError) as e:
            ctx.med(t, dev(rng_col(t, 5, 1, lo=0, hi=9)))
        assert e.value.kind == "type"


@pytest.mark.parametrize("t", [ob.U8, ob.I16, ob.I32, ob.TIME, ob.I64, ob.F64])
@pytest.mark.parametrize("n", [1, 2, 7, 16385, 1_000_003])
def test_stddev(ctx, oracle, t, n):
    col = rng_col(t, n, n + t, null_frac=0.1, lo=-100 if t != ob.U8 else 0, hi=100)
    assert close([ctx.stddev(t, dev(col))], [oracle.dev(t, col)])
    if t not in (ob.U8,):
        null = np.nan if t == ob.F64 else np.iinfo(ob.NP_OF[t]).min
        assert np.isnan(ctx.stddev(t, dev(np.full(6, null, ob.NP_OF[t]))))
        one = np.full(6, null, ob.NP_OF[t])
        one[3] = 5
        assert ctx.stddev(t, dev(one)) == 0.0


def test_stddev_exact_and_type_errors(ctx, oracle):
    x = np.array([1, 2, 3, 4, 50], np.int64)            # tests/lang.c:2599 -> 19.0263
    assert ctx.stddev(ob.I64, dev(x)) == oracle.dev(ob.I64, x)
    assert abs(ctx.stddev(ob.I64, dev(x)) - 19.0263) < 1e-4
    for t in (ob.DATE, ob.TIMESTAMP):                   # ray_sum is a type error for these, ray_dev dereferences that error
        with pytest.raises(capi.RfbError) as e:
            ctx.stddev(t, dev(rng_col(t, 5, 1, lo=0, hi=9)))
        assert e.value.kind == "type"


@pytest.mark.parametrize("n,card", [(9, 3), (20_000, 50), (300_007, 3000), (300_007, 100_000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_group_rows_and_collect(ctx, oracle, n, card, filtered):
    r = np.random.default_rng(n + card)
    keys = r.integers(0, card, n).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    gids, firsts, info = oracle.group_i64(keys, filt)
    want_rows, want_offs = oracle.group_rows(gids, info.groups, filt)
    rows, offs = ctx.group_rows(dev(gids), info.groups, dev(filt) if filtered else None)
    assert np.array_equal(host(offs), want_offs) and np.array_equal(host(rows), want_rows)
    val = rng_col(ob.F64, n, 3, null_frac=0.01)
    got = host(ctx.gather(ob.F64, dev(val), rows))      # aggr_collect = the gathered values, group by group
    assert same_f64(got, val[want_rows])


@pytest.mark.parametrize("vt", [ob.I64, ob.F64, ob.TIMESTAMP, ob.I32])
@pytest.mark.parametrize("n,card", [(9, 3), (20_000, 50), (300_007, 3000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_aggr_med(ctx, oracle, vt, n, card, filtered):
    r = np.random.default_rng(n + card + vt)
    keys = r.integers(0, card, n).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    gids, firsts, info = oracle.group_i64(keys, filt)
    val = rng_col(vt, n, vt, null_frac=0.02, lo=-1000, hi=1000)
    if vt == ob.F64:
        val[::17] = -0.0
    want, wt = oracle.aggr(ob.MED, vt, val, gids, info.groups, filt)
    got, gt = ctx.aggr(capi.A_MED, vt, dev(val), dev(gids), info.groups, dev(filt) if filtered else None)
    assert gt == wt == ob.F64
    assert same_f64(host(got), want)


@pytest.mark.parametrize("vt", [ob.I16, ob.I32, ob.TIME, ob.I64, ob.TIMESTAMP, ob.F64])
@pytest.mark.parametrize("n,card", [(9, 3), (20_000, 50), (300_007, 3000)])
@pytest.mark.parametrize("filtered", [False, True])
def test_aggr_dev(ctx, oracle, vt, n, card, filtered):
    r = np.random.default_rng(n + card + vt)
    keys = r.integers(0, card, n).astype(np.int64)
    filt = np.sort(r.choice(n, max(1, n // 3), replace=False)).astype(np.int64) if filtered else None
    gids, firsts, info = oracle.group_i64(keys, filt)
    val = rng_col(vt, n, vt, null_frac=0.02, lo=-1000, hi=1000)
    if vt == ob.F64:
        val = np.round(val * 8) / 8         # dyadic: sums of x and x*x are exact in any order -> bit-identical results
    want, wt = oracle.aggr(ob.DEV, vt, val, gids, info.groups, filt)
    got, gt = ctx.aggr(capi.A_DEV, vt, dev(val), dev(gids), info.groups, dev(filt) if filtered else None)
    assert gt == wt == ob.F64
    assert same_f64(host(got), want)


def test_aggr_dev_type_error(ctx, oracle):
    gid, val = np.zeros(4, np.int64), rng_col(ob.U8, 4, 1, lo=0, hi=9)
    with pytest.raises(ob.OracleError):
        oracle.aggr(ob.DEV, ob.U8, val, gid, 1)
    with pytest.raises(capi.RfbError) as e:
        ctx.aggr(capi.A_DEV, ob.U8, dev(val), dev(gid), 1)
    assert e.value.kind == "type"
