"""Parity at BASELINE.json's FULL sizes (1e9 rows per GPU) through size-independent properties, plus exact checks of
row windows against the CPU oracle.  The oracle cannot run 1e9 rows in seconds, so each test combines
  * algebraic identities that hold at any size (complementary predicates partition the rows and the sum; a group-by's
    sums/counts add up to the ungrouped fold; a sort's output is a sorted, stable permutation),
  * closed forms for periodic synthetic columns,
  * oracle runs on 1e6-row windows / prefixes of the very same device column.
torch is used here only to slice windows out of HBM and as an independent checker of orderings."""
import numpy as np
import pytest
import torch

from oracle import bindings as ob
from rayforce_b200 import capi

pytestmark = pytest.mark.gpu

N = 1_000_000_000
M64 = (1 << 64) - 1


def wrap(v):
    """python int -> two's-complement int64 value"""
    v &= M64
    return v - (1 << 64) if v >> 63 else v


@pytest.fixture(scope="module")
def big(ctx):
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~40 GB of free HBM")
    x = torch.empty(N, dtype=torch.int64, device="cuda")
    ctx.fill_splitmix(capi.I64, x, N, 42, 1 << 40, 0, 1009)        # every 1009th row is NULL_I64
    ctx.sync()
    yield x
    del x
    torch.cuda.empty_cache()


def test_filter_sum_1e9_partition_identity_and_oracle_windows(ctx, oracle, big):
    x, k = big, 1 << 39
    tot = ctx.fold(capi.F_ALL, capi.I64, x, N)
    lt = ctx.filter_fold(capi.LT, capi.I64, x, k, capi.F_ALL, capi.I64, x, N)
    ge = ctx.filter_fold(capi.GE, capi.I64, x, k, capi.F_ALL, capi.I64, x, N)
    fast = ctx.filter_fold(capi.LT, capi.I64, x, k, capi.F_SUM | capi.F_CNT, capi.I64, x, N)      # the headline kernel
    assert tot.rows == N and lt.rows + ge.rows == N and lt.nonnull + ge.nonnull == tot.nonnull
    assert wrap(lt.sum + ge.sum) == tot.sum
    assert min(lt.min, ge.min) == tot.min and max(lt.max, ge.max) == tot.max
    assert (fast.sum, fast.nonnull) == (lt.sum, lt.nonnull)
    assert tot.nonnull == N - N // 1009
    # i64 comparisons ignore nullness: every NULL row satisfies (< x k)
    assert lt.rows - lt.nonnull == N // 1009 and ge.rows == ge.nonnull
    # exact agreement with the oracle on row windows of the same column, and the windows add up like the whole
    acc_rows = acc_sum = 0
    for start in (0, 123_456_789, N - 1_000_003):
        w = 1_000_003
        sub = x[start:start + w]
        got = ctx.filter_fold(capi.LT, capi.I64, sub, k, capi.F_ALL, capi.I64, sub, w)
        h = sub.cpu().numpy()
        ids = oracle.where(oracle.cmp(ob.LT, ob.I64, h, ob.I64, k))
        sel = oracle.at_ids(ob.I64, h, ids)
        assert got.rows == ids.shape[0] and got.sum == int(oracle.fold(ob.SUM, ob.I64, sel)[0])
        assert got.min == int(oracle.fold(ob.MIN, ob.I64, sel)[0]) and got.max == int(oracle.fold(ob.MAX, ob.I64, sel)[0])
        acc_rows += got.rows
        acc_sum += got.sum
    assert acc_rows > 0 and acc_sum != 0


def test_host_layer_equals_device_layer_on_2e8_rows(ctx, big):
    """the e2e path (pinned host column -> chunked copies + kernels) gives the device-layer result bit for bit"""
    n = 200_000_000
    sub = big[:n]
    want = ctx.filter_fold(capi.LT, capi.I64, sub, 1 << 39, capi.F_ALL, capi.I64, sub, n)
    h = torch.empty(n, dtype=torch.int64, pin_memory=True)
    h.copy_(sub)
    got, nbytes = ctx.filter_fold_host(capi.LT, capi.I64, h.numpy(), 1 << 39, capi.F_ALL, capi.I64, h.numpy())
    assert nbytes == 8 * n
    assert (got.rows, got.nonnull, got.sum, got.min, got.max) == (want.rows, want.nonnull, want.sum, want.min, want.max)


def test_fma_avg_1e9_closed_form(ctx):
    """config 3 with integer-valued doubles: a = i % 1024, b = i % 7, c = i % 13 -> every partial sum is exact, so the
    result must equal the closed form computed with integers over the period lcm(1024, 7, 13) = 93184"""
    n = 1_000_000_000
    i = torch.arange(n, dtype=torch.int64, device="cuda")
    a, b, c = (i % 1024).double(), (i % 7).double(), (i % 13).double()
    del i
    torch.cuda.synchronize()      # the inputs were produced on torch's stream; the context launches on its own
    got = ctx.fma_fold(capi.F_SUM | capi.F_CNT, a, b, c, n)
    period = 93184
    j = np.arange(period, dtype=np.int64)
    term = (j % 1024) * (j % 7) + (j % 13)
    full, rem = divmod(n, period)
    want = int(term.sum()) * full + int(term[:rem].sum())
    assert got.nonnull == n and got.sum == float(want) and want < 2 ** 53
    assert got.avg == want / n
    del a, b, c
    torch.cuda.empty_cache()


def test_group_by_1e9_adds_up_and_first_occurrence_order(ctx, oracle, big):
    """config 4: 1e5 keys; sums/counts add up to the ungrouped fold and the group order is the oracle's on the prefix in
    which every key has appeared"""
    keys = torch.empty(N, dtype=torch.int32, device="cuda")
    ctx.fill_splitmix(capi.I32, keys, N, 7, 100_000, 0, 0)
    val = big
    gk, gs, gc = ctx.group_sum_count(capi.I32, keys, val, 100_000)
    assert gk.shape[0] == 100_000 and int(gc.sum().item()) == N
    # groups that contain a NULL value have a sticky-null sum (aggr_sum, reference core/aggr.c:1088): nearly all of them at
    # one NULL per 1009 rows; the others must add up
    hs, hc = gs.cpu().numpy(), gc.cpu().numpy()
    assert np.count_nonzero(hs == ob.NULL_I64) > 90_000
    prefix = 3_000_000
    hk, hv = keys[:prefix].cpu().numpy().astype(np.int64), val[:prefix].cpu().numpy()
    wg, wf, wi = oracle.group_i64(hk)
    assert wi.groups == 100_000, "prefix too short for every key to appear"
    assert np.array_equal(gk.cpu().numpy(), hk[wf])
    # with a predicate that rejects the NULL rows every group sum is a plain number and the totals must match the fold
    gk2, gs2, gc2 = ctx.group_sum_count(capi.I32, keys, val, 100_000, capi.GE, capi.I64, val, 0)
    tot = ctx.filter_fold(capi.GE, capi.I64, val, 0, capi.F_SUM | capi.F_CNT | capi.F_ROWS, capi.I64, val, N)
    assert int(gc2.sum().item()) == tot.rows
    assert wrap(int(gs2.cpu().numpy().astype(object).sum())) == tot.sum
    assert np.array_equal(gk2.cpu().numpy()[:1000], gk.cpu().numpy()[:1000])
    # two independent accumulate strategies (key-range partitions in shared memory vs device-wide L2 atomics) must agree bit for bit
    import os
    os.environ["RFB_GROUP_STRATEGY"] = "l2"
    try:
        gk3, gs3, gc3 = ctx.group_sum_count(capi.I32, keys, val, 100_000, capi.GE, capi.I64, val, 0)
    finally:
        del os.environ["RFB_GROUP_STRATEGY"]
    assert torch.equal(gk3, gk2) and torch.equal(gs3, gs2) and torch.equal(gc3, gc2)
    del keys
    torch.cuda.empty_cache()


@pytest.mark.parametrize("desc", [False, True])
def test_sort_2e8_is_a_stable_sorted_permutation(ctx, desc):
    n = 200_000_000
    k = torch.empty(n, dtype=torch.int64, device="cuda")
    ctx.fill_splitmix(capi.I64, k, n, 11, 1 << 20, -(1 << 19), 0)       # ~190 duplicates per key: stability matters
    perm = ctx.sort(capi.I64, k, desc)
    sk = k[perm]
    d = sk[1:] - sk[:-1]
    assert bool((d <= 0).all() if desc else (d >= 0).all()), "keys not ordered along the permutation"
    ties = d == 0
    assert bool((perm[1:][ties] > perm[:-1][ties]).all()), "equal keys must keep their original order (stable)"
    assert int(perm.sum().item()) == n * (n - 1) // 2 and int(perm.min().item()) == 0 and int(perm.max().item()) == n - 1
    chk = torch.zeros(n, dtype=torch.int8, device="cuda")
    chk[perm] = 1
    assert int(chk.sum(dtype=torch.int64).item()) == n, "not a permutation"


def test_where_and_cmp_where_1e9_large_tiles(ctx, big):
    """the compaction kernels at the size where they take their LARGE tiles (96 K rows per tile for an 8-byte predicate column, 48 K
    mask bytes): every id against torch.nonzero, on the full column"""
    k = 1 << 39
    ids = ctx.cmp_where(capi.LT, capi.I64, big, k)                  # nulls (INT64_MIN) compare below k: selected, like the reference
    want = torch.nonzero(big < k).flatten()
    assert ids.shape[0] == want.shape[0] and torch.equal(ids, want)
    del ids
    mask = (big < k).to(torch.uint8) * 7                             # any non-zero byte selects
    ids = ctx.where(mask)
    assert torch.equal(ids, want)
    del ids, want, mask
    torch.cuda.empty_cache()


@pytest.mark.parametrize("desc", [False, True])
def test_sort_wide_keys_1e8_and_narrow_window(ctx, desc):
    """full-width 64-bit keys (eight single-sweep passes over 64-bit key words) and 32-bit-window keys at an offset (four passes over
    32-bit key words, 6144-row tiles): sorted, stable, a permutation"""
    n = 100_000_000
    for modulus, shift in ((0, 0), (1 << 31, 20)):
        k = torch.empty(n, dtype=torch.int64, device="cuda")
        ctx.fill_splitmix(capi.I64, k, n, 13, modulus, 0, 0)
        if shift:
            k <<= shift
        k[::3] = k[1]                                                # a third of the rows share one key: stability
        perm = ctx.sort(capi.I64, k, desc)
        sk = k[perm]
        d_ok = (sk[1:] <= sk[:-1]) if desc else (sk[1:] >= sk[:-1])
        assert bool(d_ok.all()), "keys not ordered along the permutation"
        ties = sk[1:] == sk[:-1]
        assert bool((perm[1:][ties] > perm[:-1][ties]).all()), "equal keys must keep their original order (stable)"
        chk = torch.zeros(n, dtype=torch.int8, device="cuda")
        chk[perm] = 1
        assert int(chk.sum(dtype=torch.int64).item()) == n, "not a permutation"
        del k, perm, sk, chk, d_ok, ties
        torch.cuda.empty_cache()

