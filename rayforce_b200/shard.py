"""Multi-GPU host logic: row-range shards and the ONE exchange step of the path.

Every GPU owns rows [rank*N/P, (rank+1)*N/P) of every column (the reference splits rows the same way across its pool
threads, core/pool.c:495-507, and merges per-worker partials once: core/math.c:2222-2228, core/aggr.c:163-181).
Nothing is exchanged while scanning; the only collectives are
  * ungrouped folds : all-reduce of (rows, nonnull, sum) [SUM] and of (min, max) [MIN/MAX]; fp64 partial sums are
                      all-gathered and added in rank order so the result does not depend on the reduction tree;
  * group-by        : all-gather of each rank's (key, sum, count) rows in local first-occurrence order; the concatenation
                      (rank order) is re-grouped with the same kernels, which yields the global first-occurrence order.
The functions take a `torch.distributed` process group, so the same code runs over NCCL on the GPUs and over gloo in
the CPU tests (tests/test_shard_gloo.py, world size 2).  No column arithmetic happens here.
"""
from __future__ import annotations

from typing import Callable, Sequence, Tuple

import torch
import torch.distributed as dist

NULL_I64 = -(2 ** 63)
INF_I64 = 2 ** 63 - 1


def row_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """rows [lo, hi) owned by `rank`: contiguous, sizes differ by at most one row"""
    base, extra = divmod(n_global, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_fold_i64(rows: int, nonnull: int, total: int, mn: int, mx: int, device, group=None):
    """merge per-rank integer fold results (rfb_fold_t fields) -> (rows, nonnull, sum mod 2^64, min, max);
    min/max are NULL_I64 when no rank folded a non-null value (MINI64(NULL, y) = y, reference core/ops.h:185)"""
    s = torch.tensor([rows, nonnull, total], dtype=torch.int64, device=device)
    lo = torch.tensor([mn if nonnull else INF_I64], dtype=torch.int64, device=device)     # identity for ranks with nothing
    hi = torch.tensor([mx if nonnull else NULL_I64], dtype=torch.int64, device=device)
    dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)          # int64 SUM wraps mod 2^64 like the reference's C add
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    r, nn, tot = (int(v) for v in s.cpu())
    if nn == 0:
        return r, nn, tot, NULL_I64, NULL_I64
    return r, nn, tot, int(lo.item()), int(hi.item())


def allgather_sum_f64(partial: float, device, group=None) -> float:
    """fp64 partial sums are added in RANK ORDER on every rank: bit-identical on all ranks and independent of the
    collective's reduction tree"""
    world = dist.get_world_size(group)
    mine = torch.tensor([partial], dtype=torch.float64, device=device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    total = 0.0
    for p in parts:
        total = total + float(p.item())
    return total


def allgather_varlen(t: torch.Tensor, group=None) -> torch.Tensor:
    """concatenate 1-D tensors of different lengths from all ranks, in rank order (an all-gather-v)"""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.empty_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros(cap, dtype=t.dtype, device=t.device)
    padded[: t.shape[0]] = t
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def merge_group_partials(keys: torch.Tensor, sums: torch.Tensor, counts: torch.Tensor,
                         regroup: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], Sequence[torch.Tensor]], group=None):
    """Group-by merge.  Each rank passes its local result rows (keys in local first-occurrence order, sticky-null sums,
    counts).  They are all-gathered in rank order and re-grouped by `regroup(keys_cat, sums_cat, counts_cat)` — on the
    GPUs that is `gpu_regroup(ctx)` below (the group-index + aggregate kernels); first-occurrence numbering over the
    rank-ordered concatenation IS the global first-occurrence order because rank r holds rows before rank r+1."""
    # one size exchange and ONE data collective: the three columns travel as a [3, cap] block per rank
    world = dist.get_world_size(group)
    n = torch.tensor([keys.shape[0]], dtype=torch.int64, device=keys.device)
    sizes = torch.empty(world, dtype=torch.int64, device=keys.device)
    dist.all_gather_into_tensor(sizes, n, group=group) if keys.is_cuda else dist.all_gather(list(sizes.split(1)), n, group=group)
    sizes = [int(v) for v in sizes.cpu()]
    cap = max(max(sizes), 1)
    block = torch.zeros((3, cap), dtype=torch.int64, device=keys.device)
    block[0, : keys.shape[0]] = keys
    block[1, : sums.shape[0]] = sums
    block[2, : counts.shape[0]] = counts
    parts = torch.empty((world, 3, cap), dtype=torch.int64, device=keys.device)
    if keys.is_cuda:
        dist.all_gather_into_tensor(parts, block, group=group)
    else:
        dist.all_gather(list(parts.unbind(0)), block, group=group)
    k = torch.cat([parts[r, 0, :sz] for r, sz in enumerate(sizes)])
    s = torch.cat([parts[r, 1, :sz] for r, sz in enumerate(sizes)])
    c = torch.cat([parts[r, 2, :sz] for r, sz in enumerate(sizes)])
    return regroup(k, s, c)


def merge_group_partials_peers(ctx, keys: torch.Tensor, sums: torch.Tensor, counts: torch.Tensor, max_groups: int,
                               regroup: Callable[[torch.Tensor, torch.Tensor, torch.Tensor], Sequence[torch.Tensor]], group=None):
    """The same merge over NVLink peer memory (rfb_group_merge_peers: every rank's lists are read in place by the merge kernels,
    no collective).  `ctx.peer_groups_setup` must have bound the exchange buffers.  A key domain that is not dense is declined
    by every rank alike; they then take the all-gather + re-group route above."""
    from .capi import RfbError
    try:
        return ctx.group_merge_peers(keys, sums, counts, max_groups)
    except RfbError as e:
        if e.kind != "type":
            raise
    return merge_group_partials(keys, sums, counts, regroup, group)


def gpu_regroup(ctx):
    """regroup callback running on the GPU through the C ABI: rfb_group_i64_dev + rfb_aggr_dev(sum) + rfb_gather_dev"""
    from . import capi

    def f(k, s, c):
        gids, firsts, info = ctx.group_i64(k)
        sums, _ = ctx.aggr(capi.A_SUM, capi.I64, s, gids, info.groups)       # sticky null carries over: NULL + x = NULL
        counts, _ = ctx.aggr(capi.A_SUM, capi.I64, c, gids, info.groups)
        return ctx.gather(capi.I64, k, firsts), sums, counts
    return f
