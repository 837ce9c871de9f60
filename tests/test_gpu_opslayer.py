"""GPU parity of the reference-facing operator layer (include/rfb200_ops.h): reference obj_t objects in, objects out.
The first block replays known answers of the reference's OWN tests (tests/golden/, extracted from tests/lang.c and
tests/sort.c of the reference) through the operators `ray_sum`, `ray_lt`, `ray_add`, ... exactly as its evaluator
would call them; the second block walks the operator sequences of SURVEY §3.1-3.4 (select where / by, avg of an
expression, iasc) and compares every intermediate object with the oracle."""
import collections
import ctypes as C

import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200 import capi
from rayforce_b200.ops import Ops, OpsError, Declined, MAPFILTER, MAPGROUP
from tests import golden_io
from tests.util import rng_col, same_f64, f64_sum_ok

pytestmark = pytest.mark.gpu

NAME = {"sum": "ray_sum", "avg": "ray_avg", "min": "ray_min", "max": "ray_max", "where": "ray_where", "round": "ray_round",
        "floor": "ray_floor", "ceil": "ray_ceil", "iasc": "ray_sort_asc", "idesc": "ray_sort_desc",
        "eq": "ray_eq", "ne": "ray_ne", "lt": "ray_lt", "gt": "ray_gt", "le": "ray_le", "ge": "ray_ge", "add": "ray_add",
        "sub": "ray_sub", "mul": "ray_mul", "div": "ray_div", "fdiv": "ray_fdiv", "mod": "ray_mod", "xbar": "ray_xbar"}
UNOPS = ("round", "floor", "ceil")


@pytest.fixture(scope="module")
def ops():
    return Ops.get(0)


def test_reference_goldens_through_the_operator_layer(ops):
    """every golden whose operands the GPU layer accepts must reproduce the reference's answer; the rest must be DECLINED
    (never silently different).  Atom-only forms are the CPU body's job by contract."""
    ran, declined, bad = collections.Counter(), collections.Counter(), []
    for c in golden_io.load_cases():
        if c["op"] not in NAME:
            continue
        args = []
        for a in c["args"]:
            t, atom, arr = golden_io.decode(a)
            args.append(ops.atom(t, arr[0]) if atom else ops.vec(t, arr))
        r = ops.call(NAME[c["op"]], *args)
        try:
            got, gt = ops.value(r)
        except Declined:
            declined[c["op"]] += 1
            ops.drop(*args)
            continue
        except OpsError as e:
            if c.get("error") != e.kind:
                bad.append((c["src"], c["expr"], "raised %s" % e.kind))
            ran[c["op"]] += 1
            ops.drop(*args)
            continue
        ops.drop(*args)
        if "error" in c:
            bad.append((c["src"], c["expr"], "returned a value, reference raises " + c["error"]))
            continue
        et, eatom, earr = golden_io.decode(c["expect"])
        garr = np.atleast_1d(got)
        if et == ob.F64:
            ok = gt == et and same_f64(garr, earr, zero_sign=c["op"] not in UNOPS, max_ulp=1 if c["op"] == "fdiv" else 0)
        else:
            ok = gt == et and garr.shape == earr.shape and np.array_equal(garr.astype(np.int64), earr.astype(np.int64))
        if not ok:
            bad.append((c["src"], c["expr"], "got %d %r want %d %r" % (gt, garr[:6], et, earr[:6])))
        ran[c["op"]] += 1
    assert not bad, "%d mismatches: %r" % (len(bad), bad[:8])
    for op, least in (("sum", 8), ("min", 8), ("max", 8), ("avg", 5), ("lt", 3), ("add", 15), ("sub", 15), ("mul", 10), ("div", 50),
                      ("fdiv", 50), ("mod", 50), ("iasc", 8), ("idesc", 8), ("where", 2), ("floor", 1)):
        assert ran[op] >= least, (op, ran[op], declined[op])
    assert ops.launches > 500


def test_select_where_sum_operator_sequence(ops, oracle):
    """ray_lt -> ray_where -> filter_map -> ray_sum(MAPFILTER) and filter_collect, inside one query scope"""
    n = 300_007
    col = rng_col(ob.I64, n, 5, null_frac=0.01, lo=-1000, hi=1000)
    x, k = ops.vec(ob.I64, col), ops.atom(ob.I64, 37)
    with ops.scope():
        mask = ops.call("ray_lt", x, k)
        m, mt = ops.value(mask, drop=False)
        assert mt == ob.B8 and np.array_equal(m, oracle.cmp(ob.LT, ob.I64, col, ob.I64, 37))
        ids = ops.call("ray_where", mask)
        i, it = ops.value(ids, drop=False)
        assert it == ob.I64 and np.array_equal(i, oracle.where(m))
        lazy = ops.call("filter_map", x, ids)
        assert ops.type_of(lazy) == MAPFILTER
        sel = oracle.at_ids(ob.I64, col, i)
        for name, op in (("ray_sum", ob.SUM), ("ray_min", ob.MIN), ("ray_max", ob.MAX), ("ray_avg", ob.AVG)):
            got, gt = ops.value(ops.call(name, lazy))
            want, wt = oracle.fold(op, ob.I64, sel)
            assert gt == wt and (same_f64([got], [want]) if wt == ob.F64 else int(got) == int(want)), name
        g, gt = ops.value(ops.call("filter_collect", x, ids))
        assert gt == ob.I64 and np.array_equal(g, sel)
        ops.drop(lazy, ids, mask)
    # the fused entry point gives the same answer in one pass (outside a scope: streamed from the host column)
    got, gt = ops.value(ops.call("where_lt_sum", x, k))
    assert gt == ob.I64 and int(got) == int(oracle.fold(ob.SUM, ob.I64, sel)[0])
    got, gt = ops.value(ops.where_fold(capi.LT, 3, x, k, x))
    assert gt == ob.F64 and same_f64([got], [oracle.fold(ob.AVG, ob.I64, sel)[0]])
    ops.drop(x, k)


@pytest.mark.parametrize("filtered", [False, True])
def test_select_by_operator_sequence(ops, oracle, filtered):
    """index_group -> group_map -> ray_sum(MAPGROUP) / aggr_*; index object layout of core/index.c:1696-1699"""
    n = 200_003
    r = np.random.default_rng(1)
    keys = (r.integers(0, 1000, n) - 17).astype(np.int64)
    val = rng_col(ob.I64, n, 2, null_frac=0.001, lo=-1000, hi=1000)
    filt = np.sort(r.choice(n, n // 3, replace=False)).astype(np.int64) if filtered else None
    wg, wf, wi = oracle.group_i64(keys, filt)
    ko, vo = ops.vec(ob.I64, keys), ops.vec(ob.I64, val)
    fo = ops.vec(ob.I64, filt) if filtered else ops.NULL
    with ops.scope():
        idx = ops.call("index_group", ko, fo)
        it = ops.items(idx)
        assert ops.len_of(idx) == 7 and ops.value(it[0], drop=False)[0] == capi.INDEX_IDS
        assert int(ops.value(it[1], drop=False)[0]) == wi.groups
        assert np.array_equal(ops.value(it[2], drop=False)[0], wg)
        assert np.array_equal(ops.value(it[6], drop=False)[0], wf)
        lazy = ops.call("group_map", vo, idx)
        assert ops.type_of(lazy) == MAPGROUP
        for name, op in (("aggr_sum", ob.SUM), ("aggr_min", ob.MIN), ("aggr_max", ob.MAX), ("aggr_count", ob.COUNT), ("aggr_avg", ob.AVG)):
            got, gt = ops.value(ops.call(name, vo, idx))
            want, wt = oracle.aggr(op, ob.I64, val, wg, wi.groups, filt)
            assert gt == wt and (same_f64(got, want) if wt == ob.F64 else np.array_equal(got, want)), name
        for name, op in (("aggr_med", ob.MED), ("aggr_stddev", ob.DEV)):
            got, gt = ops.value(ops.call(name, vo, idx))
            want, wt = oracle.aggr(op, ob.I64, val, wg, wi.groups, filt)
            assert gt == wt == ob.F64 and same_f64(got, want), name
        want_rows, want_offs = oracle.group_rows(wg, wi.groups, filt)
        for name, flat in (("aggr_row", want_rows), ("aggr_collect", val[want_rows])):
            lst = ops.call(name, vo, idx)
            its = ops.items(lst)
            assert ops.type_of(lst) == 0 and len(its) == wi.groups
            for g in (0, 1, wi.groups // 2, wi.groups - 1):
                v, vt = ops.value(its[g], drop=False)
                assert vt == ob.I64 and np.array_equal(v, flat[want_offs[g]:want_offs[g + 1]]), (name, g)
            ops.drop(lst)                                       # (the builtin host drops a list's items with it)
        got, gt = ops.value(ops.call("ray_med", lazy))           # (med v) by k: the lazy pair reaches ray_med (core/math.c:2592)
        assert same_f64(got, oracle.aggr(ob.MED, ob.I64, val, wg, wi.groups, filt)[0])
        got, gt = ops.value(ops.call("ray_sum", lazy))           # FN_AGGR functions receive the lazy pair (eval.c:737)
        assert np.array_equal(got, oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
        ops.drop(lazy, idx)
    ops.drop(ko, vo)
    if filtered:
        ops.drop(fo)


def test_avg_of_expression_operator_sequence(ops, oracle):
    """(avg (+ (* a b) c)) as ray_mul -> ray_add -> ray_avg"""
    n = 250_001
    a, b, c = (rng_col(ob.F64, n, s, null_frac=0.01, lo=0, hi=1) for s in (1, 2, 3))
    ao, bo, co = (ops.vec(ob.F64, v) for v in (a, b, c))
    with ops.scope():
        t1 = ops.call("ray_mul", ao, bo)
        t2 = ops.call("ray_add", t1, co)
        want = oracle.binop(ob.ADD, ob.F64, oracle.binop(ob.MUL, ob.F64, a, ob.F64, b)[0], ob.F64, c)[0]
        assert same_f64(ops.value(t2, drop=False)[0], want)
        got, gt = ops.value(ops.call("ray_avg", t2))
        cnt = np.count_nonzero(~np.isnan(want))
        assert gt == ob.F64 and f64_sum_ok(float(got) * cnt, float(oracle.fold(ob.SUM, ob.F64, want)[0]), oracle.sum_f64_exact(want)) or \
            abs(float(got) - float(oracle.fold(ob.AVG, ob.F64, want)[0])) <= 4 * np.spacing(abs(float(got)))
        ops.drop(t1, t2)
    ops.drop(ao, bo, co)


def test_large_operands_and_results_use_the_staged_copies(ops, oracle):
    """16 MB operands / results: host objects are pageable, so both directions go through the copier-thread ring"""
    n = 2_000_003
    x = rng_col(ob.I64, n, 9, null_frac=0.01, lo=-1000, hi=1000)
    xo, k = ops.vec(ob.I64, x), ops.atom(ob.I64, 10)
    got, gt = ops.value(ops.call("ray_add", xo, k))
    assert gt == ob.I64 and np.array_equal(got, oracle.binop(ob.ADD, ob.I64, x, ob.I64, 10)[0])
    perm, pt = ops.value(ops.call("ray_sort_asc", xo))
    assert np.array_equal(perm, oracle.sort(ob.I64, x))
    ops.drop(xo, k)


def test_errors_and_declines_follow_the_reference(ops):
    a, b = ops.vec(ob.I64, np.arange(10)), ops.vec(ob.I64, np.arange(11))
    for name in ("ray_lt", "ray_add"):
        with pytest.raises(OpsError) as e:                        # vec (op) vec of different length -> "length"
            ops.value(ops.call(name, a, b))
        assert e.value.kind == "length"
    d = ops.vec(ob.DATE, np.arange(10))
    with pytest.raises(OpsError) as e:                            # (sum [2020.02.03 ...]) -> "type" (tests/lang.c:2463)
        ops.value(ops.call("ray_sum", d))
    assert e.value.kind == "type"
    u = ops.vec(ob.U8, np.arange(10))
    with pytest.raises(OpsError) as e:                            # U8 vectors do not compare (core/cmp.c:78-85)
        ops.value(ops.call("ray_eq", u, u))
    assert e.value.kind == "type"
    x, y = ops.atom(ob.I64, 1), ops.atom(ob.I64, 2)
    with pytest.raises(Declined):                                 # atom (op) atom: the CPU body's job
        ops.value(ops.call("ray_add", x, y))
    k32 = ops.vec(ob.I32, np.arange(10))
    with pytest.raises(Declined):                                 # no single-key grouping on I32 in the reference
        ops.value(ops.call("index_group", k32, ops.NULL))
    ops.drop(a, b, d, u, x, y, k32)


def test_min_rows_threshold_declines_small_inputs(ops):
    a = ops.vec(ob.I64, np.arange(100))
    ops.L.rfb_ops_set_min_rows(1000)
    try:
        with pytest.raises(Declined):
            ops.value(ops.call("ray_sum", a))
    finally:
        ops.L.rfb_ops_set_min_rows(0)
    assert int(ops.value(ops.call("ray_sum", a))[0]) == 4950
    ops.drop(a)


def test_scope_does_not_serve_a_stale_image_when_the_host_reuses_a_block(ops):
    with ops.scope():
        a = ops.vec(ob.I64, np.arange(1000))
        assert int(ops.value(ops.call("ray_sum", a))[0]) == 499500
        # same block, same shape, different content (what a freed-and-reallocated temporary looks like)
        import ctypes as C
        new = np.arange(1000, dtype=np.int64) * 3
        C.memmove(a + 16, new.ctypes.data, new.nbytes)
        assert int(ops.value(ops.call("ray_sum", a))[0]) == 499500 * 3
        ops.drop(a)


def test_lazily_materialised_results(ops, oracle):
    """EXPERIMENTAL mode (RFB200_LAZY=1 / rfb_ops_set_lazy): inside a scope large results stay on the device, their host
    pages are protected and filled on the first CPU access.  GPU consumers must find the device image without touching
    the host bytes; a CPU reader must still see exactly the reference's bytes."""
    import ctypes as C
    n = 2_000_003
    col = rng_col(ob.I64, n, 21, null_frac=0.01, lo=-1000, hi=1000)
    x, k = ops.vec(ob.I64, col), ops.atom(ob.I64, 37)
    stats = (C.c_long * 4)()
    ops.L.rfb_ops_lazy_stats(stats)
    before = list(stats)
    ops.L.rfb_ops_set_lazy.argtypes = [C.c_int, C.c_int64]
    ops.L.rfb_ops_set_lazy(1, 1 << 20)
    try:
        with ops.scope():
            mask = ops.call("ray_lt", x, k)                 # 2 MB mask: stays on the device
            ids = ops.call("ray_where", mask)                # consumes the device image of the mask; ~8 MB of ids stay too
            lazy = ops.call("filter_map", x, ids)
            got, gt = ops.value(ops.call("ray_sum", lazy))   # consumes the device image of the ids
            m_want = oracle.cmp(ob.LT, ob.I64, col, ob.I64, 37)
            i_want = oracle.where(m_want)
            assert int(got) == int(oracle.fold(ob.SUM, ob.I64, oracle.at_ids(ob.I64, col, i_want))[0])
            ops.L.rfb_ops_lazy_stats(stats)
            assert stats[0] - before[0] == 2 and stats[1] - before[1] == 0, list(stats)     # two lazy results, no fault yet
            i_got, _ = ops.value(ids, drop=False)            # a CPU reader: faults the ids in
            assert np.array_equal(i_got, i_want)
            ops.L.rfb_ops_lazy_stats(stats)
            assert stats[1] - before[1] == 1
            g2, _ = ops.value(ops.call("ray_sum", lazy))     # ids now come from the (materialised) host copy again
            assert int(g2) == int(got)
            ops.drop(lazy)
        m_got, _ = ops.value(mask)                           # never touched inside the scope: filled at scope end
        assert np.array_equal(m_got, m_want)
        ops.drop(ids)
        # a result dropped by the host before anyone touched it is simply forgotten
        with ops.scope():
            t = ops.call("ray_add", x, k)
            ops.drop(t)
        ops.L.rfb_ops_lazy_stats(stats)
        assert stats[0] - before[0] == 3
    finally:
        ops.L.rfb_ops_set_lazy(0, 0)
    ops.drop(x, k)


def test_med_dev_vectors_and_filtered(ops, oracle):
    """ray_med / ray_dev on plain vectors and on MAPFILTER pairs (the reference collects, then computes: core/math.c:2605-2608)"""
    n = 120_011
    x = rng_col(ob.I64, n, 3, null_frac=0.05, lo=-500, hi=500)
    ids = np.sort(np.random.default_rng(4).choice(n, n // 2, replace=False)).astype(np.int64)
    xo, io = ops.vec(ob.I64, x), ops.vec(ob.I64, ids)
    with ops.scope():
        got, gt = ops.value(ops.call("ray_med", xo))
        assert gt == ob.F64 and same_f64([got], [oracle.med(ob.I64, x)])
        got, gt = ops.value(ops.call("ray_dev", xo))
        assert gt == ob.F64 and abs(got - oracle.dev(ob.I64, x)) <= 1e-12 * abs(got)
        lazy = ops.call("filter_map", xo, io)
        got, gt = ops.value(ops.call("ray_med", lazy))
        assert same_f64([got], [oracle.med(ob.I64, x[ids])])
        got, gt = ops.value(ops.call("ray_dev", lazy))
        assert abs(got - oracle.dev(ob.I64, x[ids])) <= 1e-12 * abs(got)
        ops.drop(lazy)
    f = ops.vec(ob.F64, np.arange(70_000, dtype=np.float64))
    with pytest.raises(OpsError) as e:          # ray_med has no F64 vector case in the reference
        ops.value(ops.call("ray_med", f))
    assert e.value.kind == "type"
    ops.drop(xo, io, f)


@pytest.mark.parametrize("t", [ob.I64, ob.F64, ob.I32, ob.I16, ob.DATE, ob.TIMESTAMP])
@pytest.mark.parametrize("combine", [True, False])
def test_parted_aggregates(ops, oracle, t, combine):
    """aggr_sum/min/max/avg over a PARTED column (PARTED_MAP, core/aggr.c:183-260): every partition folded on
    the device, combined or returned per partition; nulls follow the GROUPED semantics (sticky sum)"""
    r = np.random.default_rng(t)
    lens = [70_001, 1, 130_000, 65_536]
    # (I16 sums run in 16 bits in the reference and a running sum that lands exactly on 0x8000 turns sticky-null: an
    #  order-dependent artefact, see test_aggr — keep 16-bit sums far from wrapping)
    span = 3 if t == ob.I16 else 500
    parts = [rng_col(t, n, t * 10 + i, null_frac=0.0 if i % 2 else 0.001, lo=-span, hi=span + 1) for i, n in enumerate(lens)]
    if t == ob.F64:
        parts = [np.round(p * 8) / 8 for p in parts]
    val = ops.parted(t, parts)
    idx = ops.parted_index(1 if combine else len(parts))
    with ops.scope():
        for name, op in (("aggr_sum", ob.SUM), ("aggr_min", ob.MIN), ("aggr_max", ob.MAX), ("aggr_avg", ob.AVG)):
            try:
                want, wt = oracle.parted_aggr(op, t, parts, combine)
            except ob.OracleError:
                with pytest.raises(Declined):            # combinations the reference's partials have no case for stay on the CPU body
                    ops.value(ops.call(name, val, idx))
                continue
            got, gt = ops.value(ops.call(name, val, idx))
            assert gt == wt, (name, gt, wt)
            assert same_f64(np.atleast_1d(got), want) if wt == ob.F64 else np.array_equal(np.atleast_1d(got), want), name
    ops.drop(val, idx)


def test_parted_aggregates_with_partition_filter(ops, oracle):
    """a parted filter (core/aggr.c:209-243): NULL entry = partition excluded, atom -1 = all its rows, I64 vector = those rows"""
    parts = [rng_col(ob.I64, n, i, null_frac=0.0, lo=-500, hi=500) for i, n in enumerate([80_000, 90_000, 100_000, 70_000])]
    ids = np.sort(np.random.default_rng(1).choice(100_000, 30_000, replace=False)).astype(np.int64)
    val = ops.parted(ob.I64, parts)
    filt = ops.list_of([ops.NULL, ops.atom(ob.I64, -1), ops.vec(ob.I64, ids), ops.vec(ob.I64, np.empty(0, np.int64))])
    import ctypes as C
    C.c_int8.from_address(filt + 2).value = 77 + ob.I64
    idx = ops.parted_index(1, filt)
    with ops.scope():
        got, gt = ops.value(ops.call("aggr_sum", val, idx))
        assert gt == ob.I64 and int(got[0]) == int(parts[1].sum() + parts[2][ids].sum())
        got, gt = ops.value(ops.call("aggr_max", val, idx))
        assert int(got[0]) == max(int(parts[1].max()), int(parts[2][ids].max()))
    ops.drop(val, idx)


# ---------------------------------------------------------------- round 2: masks, at_ids, sorted values / tables, first / last, multi-key index

def test_mask_logic_in_place_and_not(ops, oracle):
    """`and` / `or` fold into their first operand in place (core/logic.c:34-86); ray_not makes a new mask (core/order.c:422-443)"""
    n = 300_011
    r = np.random.default_rng(3)
    a, b = (r.random(n) < 0.3).astype(np.uint8), (r.random(n) < 0.6).astype(np.uint8)
    for is_or, fn in ((0, np.logical_and), (1, np.logical_or)):
        ao, bo = ops.vec(ob.B8, a), ops.vec(ob.B8, b)
        assert ops.L.rfb_mask_logic_inplace(is_or, ao, bo) == 1
        got, gt = ops.value(ao, drop=False)
        assert gt == ob.B8 and np.array_equal(got, fn(a, b).astype(np.uint8))
        assert np.array_equal(ops.value(bo, drop=False)[0], b)                   # the right operand is untouched
        ids, _ = ops.value(ops.call("ray_where", ao))                            # the updated mask is what ray_where sees
        assert np.array_equal(ids, np.flatnonzero(fn(a, b)))
        for atom in (0, 1):
            co, k = ops.vec(ob.B8, a), ops.atom(ob.B8, atom)
            assert ops.L.rfb_mask_logic_inplace(is_or, co, k) == 1
            assert np.array_equal(ops.value(co, drop=False)[0], fn(a, np.full(n, atom)).astype(np.uint8))
            ops.drop(co, k)
        short = ops.vec(ob.B8, b[:10])
        assert ops.L.rfb_mask_logic_inplace(is_or, ao, short) == 0               # length mismatch: the CPU body's type error
        ops.drop(ao, bo, short)
    ao = ops.vec(ob.B8, a)
    got, gt = ops.value(ops.call("ray_not", ao))
    assert gt == ob.B8 and np.array_equal(got, (a == 0).astype(np.uint8))
    i64v = ops.vec(ob.I64, np.arange(5))
    with pytest.raises(Declined):
        ops.value(ops.call("ray_not", i64v))
    ops.drop(ao, i64v)


@pytest.mark.parametrize("t", [ob.I64, ob.F64, ob.I32, ob.I16, ob.U8, ob.TIMESTAMP, ob.DATE])
def test_sorted_values_asc_desc(ops, oracle, t):
    """ray_asc / ray_desc (core/order.c:74-244) = x[ray_sort_asc(x)] with ATTR_ASC / ATTR_DESC set and ATTR_DISTINCT kept"""
    n = 100_003
    x = rng_col(t, n, seed=t + 40, null_frac=0.02, lo=-500 if t != ob.U8 else 0, hi=500 if t != ob.U8 else 200)
    xo = ops.vec(t, x)
    for name, desc, attr in (("ray_asc", 0, 2), ("ray_desc", 1, 4)):
        r = ops.call(name, xo)
        assert ops.attrs_of(r) == attr
        got, gt = ops.value(r)
        want = x[oracle.sort(t, x, desc)]
        assert gt == t and (same_f64(got, want) if t == ob.F64 else np.array_equal(got, want)), name
    ops.drop(xo)
    s = ops.vec(ob.SYMBOL, np.arange(10))
    with pytest.raises(Declined):                                                # symbols order by their strings: CPU body
        ops.value(ops.call("ray_asc", s))
    ops.drop(s)


def test_at_ids_vector_and_table(ops, oracle):
    """at_ids (core/rayforce.c:1100-1201): a vector, or every column of a table, gathered by a bare array of row ids"""
    n, m = 200_003, 70_001
    r = np.random.default_rng(8)
    a, b, c = rng_col(ob.I64, n, 1, null_frac=0.01, lo=-9, hi=9), rng_col(ob.F64, n, 2, null_frac=0.01, lo=0, hi=1), rng_col(ob.I16, n, 3, lo=-9, hi=9)
    ids = r.integers(0, n, m).astype(np.int64)
    ao, bo, co, io = ops.vec(ob.I64, a), ops.vec(ob.F64, b), ops.vec(ob.I16, c), ops.vec(ob.I64, ids)
    got, gt = ops.value(ops.L.rfb_at_ids(ao, io + 16, m))
    assert gt == ob.I64 and np.array_equal(got, a[ids])
    t = ops.table([11, 22, 33], [ao, bo, co])                                    # (the table owns the columns from here on)
    res = ops.L.rfb_at_ids(t, io + 16, m)
    assert ops.type_of(res) == 98
    names, cols = ops.items(res)
    assert np.array_equal(ops.value(names, drop=False)[0], [11, 22, 33])
    for col, want, wt in zip(ops.items(cols), (a, b, c), (ob.I64, ob.F64, ob.I16)):
        v, vt = ops.value(col, drop=False)
        assert vt == wt and (same_f64(v, want[ids]) if wt == ob.F64 else np.array_equal(v, want[ids]))
    ops.drop(res, t, io)


@pytest.mark.parametrize("desc", [0, 1])
def test_table_sorted_by_one_and_several_columns(ops, oracle, desc):
    """ray_xasc / ray_xdesc (core/order.c:246-420): y = symbol atom -> one stable sort; y = symbol vector -> one stable sort per
    key from the last to the first, each on the column as reordered so far"""
    n = 150_007
    r = np.random.default_rng(5 + desc)
    k1, k2 = r.integers(0, 7, n).astype(np.int64), r.integers(-50, 50, n).astype(np.int32)
    v = rng_col(ob.F64, n, 4, null_frac=0.01, lo=0, hi=1)
    t = ops.table([101, 102, 103], [ops.vec(ob.I64, k1), ops.vec(ob.I32, k2), ops.vec(ob.F64, v)])
    name = "ray_xdesc" if desc else "ray_xasc"
    one = ops.atom(ob.SYMBOL, 102)
    C.c_int8.from_address(one + 2).value = -ob.SYMBOL
    res = ops.call(name, t, one)
    perm = oracle.sort(ob.I32, k2, desc)
    for col, want in zip(ops.items(ops.items(res)[1]), (k1, k2, v)):
        got = ops.value(col, drop=False)[0]
        assert same_f64(got, want[perm]) if want.dtype == np.float64 else np.array_equal(got, want[perm])
    ops.drop(res)
    by = ops.vec(ob.SYMBOL, np.array([101, 102], np.int64))
    res = ops.call(name, t, by)
    perm = np.arange(n)
    for col, ct in ((k2, ob.I32), (k1, ob.I64)):                                 # last key first, stable
        perm = perm[oracle.sort(ct, col[perm], desc)]
    for col, want in zip(ops.items(ops.items(res)[1]), (k1, k2, v)):
        got = ops.value(col, drop=False)[0]
        assert same_f64(got, want[perm]) if want.dtype == np.float64 else np.array_equal(got, want[perm])
    missing = ops.atom(ob.SYMBOL, 999)
    C.c_int8.from_address(missing + 2).value = -ob.SYMBOL
    with pytest.raises(Declined):                                                # unknown column: the CPU body's error
        ops.value(ops.call(name, t, missing))
    ops.drop(res, by, one, missing, t)


@pytest.mark.parametrize("filtered", [False, True])
def test_aggr_first_last_and_multi_key_index(ops, oracle, filtered):
    n = 120_011
    r = np.random.default_rng(17)
    ka, kb = r.integers(0, 40, n).astype(np.int64), (r.integers(0, 25, n) * 1000).astype(np.int64)
    val = rng_col(ob.I64, n, 6, null_frac=0.3, lo=-100, hi=100)
    filt = np.sort(r.choice(n, n // 3, replace=False)).astype(np.int64) if filtered else None
    wg, wf, groups = oracle.group_multi([ka, kb], filt)
    keys = ops.list_of([ops.vec(ob.I64, ka), ops.vec(ob.I64, kb)])
    vo = ops.vec(ob.I64, val)
    fo = ops.vec(ob.I64, filt) if filtered else ops.NULL
    with ops.scope():
        idx = ops.call("index_group_list", keys, fo)
        it = ops.items(idx)
        assert ops.len_of(idx) == 7 and int(ops.value(it[1], drop=False)[0]) == groups
        assert np.array_equal(ops.value(it[2], drop=False)[0], wg) and np.array_equal(ops.value(it[6], drop=False)[0], wf)
        got, gt = ops.value(ops.call("aggr_sum", vo, idx))
        assert np.array_equal(got, oracle.aggr(ob.SUM, ob.I64, val, wg, groups, filt)[0])
        got, gt = ops.value(ops.call("aggr_first", vo, idx))
        assert gt == ob.I64 and np.array_equal(got, oracle.aggr(ob.FIRST, ob.I64, val, wg, groups, filt)[0])
        got, gt = ops.value(ops.call("aggr_last", vo, idx))                      # builtin host: one executor -> one chunk
        assert gt == ob.I64 and np.array_equal(got, oracle.aggr(ob.LAST, ob.I64, val, wg, groups, filt)[0])
        ops.drop(idx)
    ops.drop(keys, vo)
    if filtered:
        ops.drop(fo)


@pytest.mark.parametrize("n,card", [(1000, 37), (30_000, 5000), (400_003, 60_000)])
def test_distinct_sparse_range_keeps_the_reference_slot_order(ops, oracle, n, card):
    """ray_distinct's hash branch (core/index.c:579-603): the distinct keys in the SLOT order of the reference's open-addressing
    table (next_prime(ceil(len / 0.75)) slots, key % size, linear probing, rows inserted in row order)"""
    r = np.random.default_rng(n + card)
    pool = r.integers(0, 1 << 62, card).astype(np.int64)
    keys = pool[r.integers(0, card, n)]
    keys[0], keys[-1] = 0, (1 << 62) + 12345
    ko = ops.vec(ob.I64, keys)
    res = ops.call("ray_distinct", ko)
    assert ops.attrs_of(res) & 1
    got, gt = ops.value(res)
    assert gt == ob.I64 and np.array_equal(got, oracle.distinct(keys))
    neg = keys.copy()
    neg[5] = -7                                                                  # negative keys: the CPU body's (out-of-table) business
    no = ops.vec(ob.I64, neg)
    with pytest.raises(Declined):
        ops.value(ops.call("ray_distinct", no))
    ops.drop(ko, no)
