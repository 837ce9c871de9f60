#!/usr/bin/env python
"""Parity of the sharded path on N GPUs (run under torchrun, NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py
Rank r holds rows [r*n, (r+1)*n) of the same synthetic columns; the merged filter+fold and filter+group-by+sum results
must equal what ONE GPU computes over all N*n rows (rank 0 recomputes that when it fits) and the CPU oracle at reduced n."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayforce_b200 import Context, capi, shard  # noqa: E402

GOLDEN = 0x9E3779B97F4A7C15


def sseed(seed, first_row):
    return (seed + first_row * GOLDEN) & 0xFFFFFFFFFFFFFFFF


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20_000_000, help="rows per GPU")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    st = torch.cuda.Stream()
    ctx = Context(local, stream=st.cuda_stream)
    n = args.rows
    K, KV = 1 << 39, 1 << 19

    def cols(first_row, rows):
        with torch.cuda.stream(st):
            x = torch.empty(rows, dtype=torch.int64, device=dev)
            k = torch.empty(rows, dtype=torch.int64, device=dev)
            v = torch.empty(rows, dtype=torch.int64, device=dev)
        ctx.fill_splitmix(capi.I64, x, rows, sseed(42, first_row), 1 << 40, 0, 0)
        ctx.fill_splitmix(capi.I64, k, rows, sseed(7, first_row), 100_000, 0, 0)
        ctx.fill_splitmix(capi.I64, v, rows, sseed(9, first_row), 1 << 20, 0, 0)
        ctx.sync()
        return x, k, v

    x, k, v = cols(rank * n, n)
    with torch.cuda.stream(st):
        # --- filter + fold, merged with one all-reduce
        r = ctx.filter_fold(capi.LT, capi.I64, x, K, capi.F_ALL, capi.I64, x, n)
        merged = shard.allreduce_fold_i64(r.rows, r.nonnull, r.sum, r.min, r.max, dev)
        # --- the same merge as ONE tiny kernel over NVLink peer mailboxes (rfb_fold_allreduce_peers), three times in a row
        peer_ok = True
        try:
            ctx.peer_mailbox_setup(rank, world)
            for _ in range(3):
                ctx.filter_fold_async(capi.LT, capi.I64, x, K, capi.F_ALL, capi.I64, x, n)
                pr = ctx.fold_allreduce_peers(capi.I64)
                peer_ok &= (pr.rows, pr.nonnull, pr.sum, pr.min, pr.max) == tuple(merged)
            # queued back to back without a host synchronisation (what bench.py's device-resident loop does): seven fold +
            # exchange pairs, alternating predicates so that a stale mailbox buffer would show, one collection at the end
            for i in range(7):
                ctx.filter_fold_async(capi.LT if i % 2 == 0 else capi.GE, capi.I64, x, K, capi.F_ALL, capi.I64, x, n)
                ctx.fold_allreduce_peers_async(capi.I64)
            pr = ctx.fold_peers_result(capi.I64)
            peer_ok &= (pr.rows, pr.nonnull, pr.sum, pr.min, pr.max) == tuple(merged)
        except Exception as e:                                   # CUDA IPC not available in this sandbox: reported, not fatal
            peer_ok = None
            peer_err = str(e)
        # --- filter + group-by + sum/count, merged with one all-gather-v and a re-group on every rank
        lk, ls, lc = ctx.group_sum_count(capi.I64, k, v, 100_000, capi.LT, capi.I64, v, KV)
        mk, ms, mc = shard.merge_group_partials(lk, ls, lc, shard.gpu_regroup(ctx))
        # --- the same merge over NVLink peer memory (rfb_group_merge_peers), three times in a row (both halves of the exchange
        #     buffers and their reuse), plus a list with null sums and keys only some ranks hold
        peer_group_ok = True
        try:
            ctx.peer_groups_setup(rank, world, 1 << 18)
            for _ in range(3):
                pk, ps, pc = ctx.group_merge_peers(lk, ls, lc, 1 << 18)
                peer_group_ok &= bool(torch.equal(pk, mk) and torch.equal(ps, ms) and torch.equal(pc, mc))
            ok2 = torch.arange(1000 * rank, 1000 * rank + 3000, dtype=torch.int64, device=dev).flip(0).contiguous()     # overlapping key ranges
            os2 = torch.full_like(ok2, 7 + rank)
            os2[::5] = capi.NULL_I64 if rank % 2 == 0 else 3
            oc2 = torch.full_like(ok2, 2)
            pk, ps, pc = ctx.group_merge_peers(ok2, os2, oc2, 1 << 18)
            wk, ws, wc = shard.merge_group_partials(ok2, os2, oc2, shard.gpu_regroup(ctx))
            peer_group_ok &= bool(torch.equal(pk, wk) and torch.equal(ps, ws) and torch.equal(pc, wc))
            wide = torch.tensor([0, 1 << 40], dtype=torch.int64, device=dev)                                               # not a dense domain: declined by all ranks
            pk, ps, pc = shard.merge_group_partials_peers(ctx, wide, wide, wide, 1 << 18, shard.gpu_regroup(ctx))
            peer_group_ok &= int(pk.shape[0]) == 2 and int(pc[1].item()) == world * (1 << 40)
        except Exception as e:
            peer_group_ok = None
            peer_group_err = str(e)
        torch.cuda.synchronize()
    ok = True
    report = {"world": world, "rows_per_gpu": n, "merged_fold": list(merged), "groups": int(mk.shape[0]), "peer_mailbox_allreduce_equal": peer_ok}
    report["peer_group_merge_equal"] = peer_group_ok
    if peer_ok is None:
        report["peer_mailbox_error"] = peer_err
    if peer_group_ok is None:
        report["peer_group_merge_error"] = peer_group_err
    ok &= peer_ok is not False and peer_group_ok is not False
    # every rank must hold the same merged result
    sig = torch.tensor([merged[2], int(ms.sum().item()), int(mc.sum().item()), int(mk[:100].sum().item())], dtype=torch.int64, device=dev)
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig)
    ok &= all(torch.equal(s, sigs[0]) for s in sigs)
    if rank == 0:
        del x, k, v
        torch.cuda.empty_cache()
        X, Kc, V = cols(0, n * world)                       # the same rows on ONE GPU
        with torch.cuda.stream(st):
            r1 = ctx.filter_fold(capi.LT, capi.I64, X, K, capi.F_ALL, capi.I64, X, n * world)
            k1, s1, c1 = ctx.group_sum_count(capi.I64, Kc, V, 100_000, capi.LT, capi.I64, V, KV)
            torch.cuda.synchronize()
        single = (r1.rows, r1.nonnull, r1.sum, r1.min, r1.max)
        ok &= tuple(merged) == single
        ok &= torch.equal(mk, k1) and torch.equal(ms, s1) and torch.equal(mc, c1)
        report.update(single_gpu_fold=list(single), group_rows_equal=bool(torch.equal(mk, k1)), ok=bool(ok))
        print(json.dumps(report), flush=True)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
