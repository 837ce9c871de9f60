"""rfb_mgpu_*: every visible GPU from one host process behind the C ABI.  Host columns are cut into row ranges, one
per device; the partial folds / group lists are merged on the host.  The merged result must equal the oracle's over all rows —
on a 1-GPU box the same code runs with one shard, on an N-GPU box (gpurun --gpus N) with N."""
import numpy as np
import pytest

from oracle import bindings as ob
from rayforce_b200 import MultiGpu, capi
from tests.util import rng_col, f64_sum_ok

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mg():
    m = MultiGpu(0)
    yield m
    m.close()


@pytest.mark.parametrize("n", [0, 1, 1000, 1_000_003, 5_000_011])
def test_filter_fold_over_all_devices(mg, oracle, n):
    x = rng_col(ob.I64, n, seed=n + 1, null_frac=0.01, lo=-(1 << 40), hi=1 << 40)
    k = 12345
    got, nbytes = mg.filter_fold_host(capi.LT, capi.I64, x, k, capi.F_ALL, capi.I64, x, chunk_rows=1 << 17)
    ids = oracle.where(oracle.cmp(ob.LT, ob.I64, x, ob.I64, k))
    sel = oracle.at_ids(ob.I64, x, ids)
    assert got.rows == ids.shape[0] and nbytes == n * 8
    if sel.shape[0] and np.count_nonzero(sel != ob.NULL_I64):
        assert got.sum == int(oracle.fold(ob.SUM, ob.I64, sel)[0])
        assert got.min == int(oracle.fold(ob.MIN, ob.I64, sel)[0]) and got.max == int(oracle.fold(ob.MAX, ob.I64, sel)[0])
    # plain fold of an F64 column: error-free partial sums merge to within 1 ULP of the exact sum
    f = rng_col(ob.F64, n, seed=n + 2, null_frac=0.01, lo=-1e6, hi=1e6)
    got, _ = mg.filter_fold_host(None, None, None, None, capi.F_ALL, capi.F64, f)
    if n:
        assert f64_sum_ok(got.sum, float(oracle.fold(ob.SUM, ob.F64, f)[0]), oracle.sum_f64_exact(f))


@pytest.mark.parametrize("with_pred", [False, True])
@pytest.mark.parametrize("key_type", [ob.I32, ob.I64])
def test_group_sum_count_over_all_devices(mg, oracle, key_type, with_pred):
    n, card = 3_000_017, 50_000
    r = np.random.default_rng(card + key_type)
    keys64 = (r.integers(0, card, n) - 1000).astype(np.int64)
    val = r.integers(0, 1 << 20, n).astype(np.int64)
    val[r.random(n) < 0.0005] = ob.NULL_I64
    keys = keys64.astype(ob.NP_OF[key_type])
    if with_pred:
        filt = oracle.where(oracle.cmp(ob.LT, ob.I64, val, ob.I64, 1 << 19))
        gk, gs, gc, nb = mg.group_sum_count_host(key_type, keys, val, card + 8, cmp_op=capi.LT, pred_type=ob.I64, pred=val, k=1 << 19)
    else:
        filt = None
        gk, gs, gc, nb = mg.group_sum_count_host(key_type, keys, val, card + 8)
    wg, wf, wi = oracle.group_i64(keys64, filt)
    rows = wf if filt is None else filt[wf]
    assert np.array_equal(gk, keys64[rows])                                      # first-occurrence order over ALL rows
    assert np.array_equal(gs, oracle.aggr(ob.SUM, ob.I64, val, wg, wi.groups, filt)[0])
    assert np.array_equal(gc, oracle.aggr(ob.COUNT, ob.I64, val, wg, wi.groups, filt)[0])
    assert nb == n * (np.dtype(ob.NP_OF[key_type]).itemsize + 8)
