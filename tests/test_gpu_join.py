"""GPU parity for the equi-join row matching (k_join.cu) against the CPU oracle (pinned against the compiled
reference's ray_find / index_left_join_obj / index_inner_join_obj in tests/test_oracle_vs_reference.py): bit-exact row ids."""
import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200.ops import Ops, Declined
from tests.util import dev, host

pytestmark = pytest.mark.gpu


def tables(ncols, nb, np_, seed, wide=False):
    r = np.random.default_rng(seed)
    scale = (1 << 50) if wide else 1
    bcols = [(r.integers(-6, 12 + 3 * c, nb) * scale).astype(np.int64) for c in range(ncols)]
    pcols = [(r.integers(-8, 14 + 3 * c, np_) * scale).astype(np.int64) for c in range(ncols)]
    if nb > 5:
        bcols[-1][::7] = ob.NULL_I64
    if np_ > 5:
        pcols[-1][::5] = ob.NULL_I64
    return bcols, pcols


@pytest.mark.parametrize("ncols", [1, 2, 5])
@pytest.mark.parametrize("nb,np_", [(1, 1), (3, 1000), (1000, 3), (70_001, 200_003), (500_000, 100_000)])
@pytest.mark.parametrize("wide", [False, True])
def test_find_rows_and_inner_join(ctx, oracle, ncols, nb, np_, wide):
    bcols, pcols = tables(ncols, nb, np_, ncols * 1000 + nb, wide)
    want = oracle.find_rows(bcols, pcols)
    db, dp = [dev(c) for c in bcols], [dev(c) for c in pcols]
    assert np.array_equal(host(ctx.find_rows(db, dp)), want)
    pi, bi = ctx.inner_join(db, dp)
    wpi, wbi = oracle.inner_join(bcols, pcols)
    assert np.array_equal(host(pi), wpi) and np.array_equal(host(bi), wbi)


def test_find_rows_high_cardinality_unique_keys(ctx, oracle):
    """the published join shape: (nearly) unique keys on both sides (reference docs benchmarks/inner-join.md: ij [id1 id2])"""
    n = 1_000_003
    r = np.random.default_rng(9)
    b1, b2 = r.permutation(n).astype(np.int64), r.integers(0, 1000, n).astype(np.int64)
    sel = r.integers(0, n, n)
    p1, p2 = b1[sel].copy(), b2[sel].copy()
    p2[::3] += 1                                     # a third of the probe rows have no partner
    want = oracle.find_rows([b1, b2], [p1, p2])
    got = host(ctx.find_rows([dev(b1), dev(b2)], [dev(p1), dev(p2)]))
    assert np.array_equal(got, want) and (got == ob.NULL_I64).sum() > n // 4


def test_find_rows_empty_sides(ctx):
    e, x = dev(np.empty(0, np.int64)), dev(np.arange(5, dtype=np.int64))
    assert host(ctx.find_rows([e], [x])).tolist() == [ob.NULL_I64] * 5
    assert ctx.find_rows([x], [e]).shape[0] == 0
    pi, bi = ctx.inner_join([e], [x])
    assert pi.shape[0] == 0 and bi.shape[0] == 0


def test_join_through_the_operator_layer(oracle):
    """index_left_join_obj / index_inner_join_obj / ray_find with the reference's object layout (lists of key columns)"""
    ops = Ops.get(0)
    bcols, pcols = tables(2, 90_001, 150_003, 5)
    want = oracle.find_rows(bcols, pcols)
    lo = ops.list_of([ops.vec(ob.I64, c) for c in pcols])
    ro = ops.list_of([ops.vec(ob.I64, c) for c in bcols])
    with ops.scope():
        got, gt = ops.value(ops.call("index_left_join_obj", lo, ro, 2))
        assert gt == ob.I64 and np.array_equal(got, want)
        pair = ops.call("index_inner_join_obj", lo, ro, 2)
        its = ops.items(pair)
        wpi, wbi = oracle.inner_join(bcols, pcols)
        assert np.array_equal(ops.value(its[0], drop=False)[0], wpi) and np.array_equal(ops.value(its[1], drop=False)[0], wbi)
        ops.drop(pair)
        x, y = ops.vec(ob.I64, bcols[0]), ops.vec(ob.I64, pcols[0])
        got, gt = ops.value(ops.call("ray_find", x, y))
        assert np.array_equal(got, oracle.find_rows([bcols[0]], [pcols[0]]))
        got, gt = ops.value(ops.call("ray_in", y, x))          # mask of the probe keys that occur in the build column
        assert gt == ob.B8 and np.array_equal(got, (oracle.find_rows([bcols[0]], [pcols[0]]) != ob.NULL_I64).astype(np.uint8))
        f = ops.vec(ob.F64, np.arange(70_000, dtype=np.float64))
        with pytest.raises(Declined):                     # other key types stay on the reference's ray_find
            ops.value(ops.call("ray_find", f, f))
        ops.drop(x, y, f)
    ops.drop(lo, ro)


def asof_tables(ncols, nb, np_, seed, tt=ob.I64):
    r = np.random.default_rng(seed)
    bcols = [r.integers(0, 60 + c, nb).astype(np.int64) for c in range(ncols)]
    pcols = [r.integers(0, 70 + c, np_).astype(np.int64) for c in range(ncols)]
    bt = np.sort(r.integers(0, 10_000_000, nb)).astype(ob.NP_OF[tt])
    pt = r.integers(-10, 10_000_100, np_).astype(ob.NP_OF[tt])
    return bcols, bt, pcols, pt


@pytest.mark.parametrize("ncols", [1, 2, 3])
@pytest.mark.parametrize("tt", [ob.I64, ob.I32, ob.TIMESTAMP])
@pytest.mark.parametrize("nb,np_", [(1, 5), (20, 5000), (300_007, 500_003)])
def test_asof_join(ctx, oracle, ncols, tt, nb, np_):
    """index_asof_join_obj (core/index.c:3194-3268): last build row of the probe row's key with time <= the probe time"""
    bcols, bt, pcols, pt = asof_tables(ncols, nb, np_, nb + ncols, tt)
    want = oracle.asof_join(bcols, tt, bt, pcols, pt)
    got = ctx.asof_join([dev(c) for c in bcols], tt, dev(bt), [dev(c) for c in pcols], dev(pt))
    assert np.array_equal(host(got), want)


def test_asof_join_through_the_operator_layer(oracle):
    ops = Ops.get(0)
    bcols, bt, pcols, pt = asof_tables(2, 120_001, 200_003, 3)
    lo, ro = ops.list_of([ops.vec(ob.I64, c) for c in pcols]), ops.list_of([ops.vec(ob.I64, c) for c in bcols])
    lx, rx = ops.vec(ob.I64, pt), ops.vec(ob.I64, bt)
    with ops.scope():
        got, gt = ops.value(ops.L.rfb_index_asof_join_obj(lo, lx, ro, rx))
        assert gt == ob.I64 and np.array_equal(got, oracle.asof_join(bcols, ob.I64, bt, pcols, pt))
    ops.drop(lo, ro, lx, rx)


@pytest.mark.parametrize("n,card,kmin", [(1, 1, 5), (1000, 40, -20), (300_003, 5000, -2500), (50_000, 900_000, 17), (400_001, 100_000, 0)])
def test_distinct_dense(ctx, oracle, n, card, kmin):
    keys = (np.random.default_rng(n).integers(0, card, n) + kmin).astype(np.int64)
    assert np.array_equal(host(ctx.distinct(dev(keys))), oracle.distinct(keys))
    assert np.array_equal(host(ctx.distinct(dev(np.concatenate([[0], keys]))[1:])), oracle.distinct(keys))     # unaligned column


def test_distinct_sparse_range(ctx, oracle):
    """the device entry point only takes dense ranges (RFB_ERR_ARG otherwise); the operator layer serves a sparse range by replaying
    the reference's table slot order over the distinct keys (all 10 000 distinct here: declined, that IS the reference's work)"""
    from rayforce_b200 import capi
    keys = (np.random.default_rng(1).integers(0, 1 << 50, 10_000)).astype(np.int64)
    with pytest.raises(capi.RfbError) as e:
        ctx.distinct(dev(keys))
    assert e.value.kind == "arg"
    ops = Ops.get(0)
    x = ops.vec(ob.I64, keys)
    with pytest.raises(Declined):
        ops.value(ops.call("ray_distinct", x))
    rep = np.concatenate([keys[:500]] * 20)                                          # 500 distinct keys over a 2^50 range
    y = ops.vec(ob.I64, rep)
    got, gt = ops.value(ops.call("ray_distinct", y))
    assert gt == ob.I64 and np.array_equal(got, oracle.distinct(rep))
    d = ops.vec(ob.I64, keys % 1000)
    got, gt = ops.value(ops.call("ray_distinct", d))
    assert gt == ob.I64 and np.array_equal(got, oracle.distinct(keys % 1000))
    ops.drop(x, y, d)


@pytest.mark.parametrize("ncols", [1, 2])
@pytest.mark.parametrize("vt", [ob.I64, ob.F64])
@pytest.mark.parametrize("nl,nr,span", [(1, 1, 10), (300, 50, 5_000), (50_000, 400_000, 10_000_000), (200_003, 100_000, 1_000_000)])
def test_window_join(ctx, oracle, ncols, vt, nl, nr, span):
    """window-join / window-join1 aggregates (core/join.c:358-485, core/index.c:3287-3346, core/aggr.c:131-160) against the oracle,
    which is pinned against the reference's own window-join through Rayfall: keys missing on either side, windows without rows,
    blocks of one row, nulls in the value column (sticky sum)"""
    from rayforce_b200 import capi
    from tests.util import rng_col, same_f64
    r = np.random.default_rng(nl + nr + ncols)
    lcols = [r.integers(0, 30 + c, nl).astype(np.int64) for c in range(ncols)]
    rcols = [r.integers(0, 33 + c, nr).astype(np.int64) for c in range(ncols)]
    lt = np.sort(r.integers(0, span, nl)).astype(np.int32)
    rt = r.integers(0, span, nr).astype(np.int32)
    val = rng_col(vt, nr, 7, null_frac=0.01, lo=-50, hi=50)
    if vt == ob.F64:
        val = np.round(val * 4) / 4
    order = np.lexsort([rt] + rcols[::-1])           # the right table ordered by (key tuple, time), as ray_window_join's xasc leaves it
    rcols, rt, val = [c[order] for c in rcols], rt[order], val[order]
    wlo, whi = (lt - span // 50).astype(np.int32), (lt + span // 40).astype(np.int32)
    d = dict(r=[dev(c) for c in rcols], rt=dev(rt), l=[dev(c) for c in lcols], lo=dev(wlo), hi=dev(whi), v=dev(val))
    for jt in (0, 1):
        for op, aop in ((ob.SUM, capi.A_SUM), (ob.MIN, capi.A_MIN), (ob.MAX, capi.A_MAX), (ob.COUNT, capi.A_COUNT), (ob.AVG, capi.A_AVG)):
            want, wt = oracle.window_aggr(op, vt, val, rcols, rt, lcols, wlo, whi, jt)
            got, gt = ctx.window_join(aop, vt, d["v"], d["r"], d["rt"], d["l"], d["lo"], d["hi"], jt)
            assert gt == wt, (op, jt)
            assert same_f64(host(got), want) if wt == ob.F64 else np.array_equal(host(got), want), (op, jt)
