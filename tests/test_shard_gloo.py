"""Multi-GPU host logic on CPU: world size 2 over gloo (the N>1 path of SURVEY §8e).  Each rank computes its shard's
partial result with the CPU oracle (the GPU kernels need a device), then the product's exchange step
(rayforce_b200/shard.py) merges them; the merged result must equal the oracle on the whole column."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import bindings as ob  # noqa: E402
from rayforce_b200 import shard  # noqa: E402

WORLD = 2


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def column(n):
    r = np.random.default_rng(7)
    col = r.integers(-(1 << 40), 1 << 40, n).astype(np.int64)
    col[r.random(n) < 0.02] = ob.NULL_I64
    keys = r.integers(0, 500, n).astype(np.int64)
    val = r.integers(0, 1 << 20, n).astype(np.int64)
    val[r.random(n) < 0.001] = ob.NULL_I64
    f = r.uniform(-1, 1, n)
    return col, keys, val, f


def oracle_regroup(O):
    def f(k, s, c):
        k, s, c = k.numpy(), s.numpy(), c.numpy()
        gids, firsts, info = O.group_i64(k)
        return (torch.from_numpy(k[firsts]), torch.from_numpy(O.aggr(ob.SUM, ob.I64, s, gids, info.groups)[0]),
                torch.from_numpy(O.aggr(ob.SUM, ob.I64, c, gids, info.groups)[0]))
    return f


def worker(rank, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    O = ob.Oracle()
    col, keys, val, f = column(n)
    lo, hi = shard.row_range(n, rank, WORLD)
    k = 12345
    # --- filter + fold over this rank's rows
    ids = O.where(O.cmp(ob.LT, ob.I64, col[lo:hi], ob.I64, k))
    sel = O.at_ids(ob.I64, col[lo:hi], ids)
    nn = int(O.fold(ob.CNT, ob.I64, sel)[0])
    res = shard.allreduce_fold_i64(ids.shape[0], nn, int(O.fold(ob.SUM, ob.I64, sel)[0]), int(O.fold(ob.MIN, ob.I64, sel)[0]),
                                   int(O.fold(ob.MAX, ob.I64, sel)[0]), "cpu")
    # --- a rank with nothing selected must not poison min/max
    empty = shard.allreduce_fold_i64(0, 0, 0, ob.NULL_I64, ob.NULL_I64, "cpu") if rank == 0 else \
        shard.allreduce_fold_i64(3, 2, 30, 10, 20, "cpu")
    fsum = shard.allgather_sum_f64(float(O.fold(ob.SUM, ob.F64, f[lo:hi])[0]), "cpu")
    # --- group-by sum/count of this rank's rows, then the merge
    gids, firsts, info = O.group_i64(keys[lo:hi])
    lk = keys[lo:hi][firsts]
    ls = O.aggr(ob.SUM, ob.I64, val[lo:hi], gids, info.groups)[0]
    lc = O.aggr(ob.COUNT, ob.I64, val[lo:hi], gids, info.groups)[0]
    mk, ms, mc = shard.merge_group_partials(torch.from_numpy(lk), torch.from_numpy(ls), torch.from_numpy(lc), oracle_regroup(O))
    # the peer-memory route declines a key domain that is not dense (every rank alike) and lands on the same all-gather merge
    class Declines:
        def group_merge_peers(self, *a):
            from rayforce_b200 import capi
            raise capi.RfbError(capi.ERR_TYPE, "not a dense domain")
    pk, ps, pc = shard.merge_group_partials_peers(Declines(), torch.from_numpy(lk), torch.from_numpy(ls), torch.from_numpy(lc), 1 << 10, oracle_regroup(O))
    assert torch.equal(pk, mk) and torch.equal(ps, ms) and torch.equal(pc, mc)
    var = shard.allgather_varlen(torch.arange(rank * 3 + 1, dtype=torch.int64))
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), res=np.array(res, dtype=object), empty=np.array(empty, dtype=object), fsum=fsum,
             mk=mk.numpy(), ms=ms.numpy(), mc=mc.numpy(), var=var.numpy())
    dist.destroy_process_group()


def test_row_ranges_partition_the_rows():
    for n in (0, 1, 7, 1000, 1_000_000_007):
        for w in (1, 2, 3, 8):
            spans = [shard.row_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


@pytest.mark.timeout(300)
def test_two_rank_merge_equals_single_stream_oracle(tmp_path, oracle):
    n = 200_001
    port = free_port()
    mp.spawn(worker, args=(port, n, str(tmp_path)), nprocs=WORLD, join=True)
    col, keys, val, f = column(n)
    ids = oracle.where(oracle.cmp(ob.LT, ob.I64, col, ob.I64, 12345))
    sel = oracle.at_ids(ob.I64, col, ids)
    want = (ids.shape[0], int(oracle.fold(ob.CNT, ob.I64, sel)[0]), int(oracle.fold(ob.SUM, ob.I64, sel)[0]),
            int(oracle.fold(ob.MIN, ob.I64, sel)[0]), int(oracle.fold(ob.MAX, ob.I64, sel)[0]))
    gids, firsts, info = oracle.group_i64(keys)
    outs = [np.load(os.path.join(tmp_path, "r%d.npz" % r), allow_pickle=True) for r in range(WORLD)]
    lo, hi = shard.row_range(n, 0, WORLD)
    fs = float(oracle.fold(ob.SUM, ob.F64, f[lo:hi])[0]) + float(oracle.fold(ob.SUM, ob.F64, f[hi:])[0])
    for o in outs:
        assert tuple(int(v) for v in o["res"]) == want
        assert tuple(int(v) for v in o["empty"]) == (3, 2, 30, 10, 20)
        assert float(o["fsum"]) == fs                                                # rank-order sum: identical on every rank
        assert np.array_equal(o["mk"], keys[firsts])                                 # global first-occurrence order
        assert np.array_equal(o["ms"], oracle.aggr(ob.SUM, ob.I64, val, gids, info.groups)[0])
        assert np.array_equal(o["mc"], oracle.aggr(ob.COUNT, ob.I64, val, gids, info.groups)[0])
        assert o["var"].tolist() == [0, 0, 1, 2, 3]
