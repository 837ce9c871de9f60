#!/usr/bin/env python
"""BASELINE config 5: N x 1e9 rows sharded by row range over N GPUs, `select {s: (sum v) c: (count v) from t by k where (< v c)}`
with 1e5 int32 keys; every GPU groups its shard (fused group-by, key-range partitions in shared memory), the per-GPU result rows
are all-gathered (NCCL) and re-grouped on every rank.  Weak scaling: fixed rows per GPU.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 tools/groupby_scaling.py
    python tools/groupby_scaling.py            (N = 1)
Timing: CUDA events on the stream all kernels and collectives run on, barrier + synchronize on both sides, max over ranks."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayforce_b200 import Context, capi, shard  # noqa: E402

GOLDEN = 0x9E3779B97F4A7C15


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    st = torch.cuda.Stream()
    ctx = Context(local, stream=st.cuda_stream)
    n, first = a.rows, rank * a.rows
    with torch.cuda.stream(st):
        k = torch.empty(n, dtype=torch.int32, device=dev)
        v = torch.empty(n, dtype=torch.int64, device=dev)
    ctx.fill_splitmix(capi.I32, k, n, (7 + first * GOLDEN) & (2**64 - 1), 100_000, 0, 0)
    ctx.fill_splitmix(capi.I64, v, n, (9 + first * GOLDEN) & (2**64 - 1), 1 << 20, 0, 0)
    ctx.sync()
    regroup = shard.gpu_regroup(ctx)

    def step():
        lk, ls, lc = ctx.group_sum_count(capi.I32, k, v, 100_000, capi.LT, capi.I64, v, 1 << 19)
        if world > 1:
            return shard.merge_group_partials(lk, ls, lc, regroup)
        return lk, ls, lc

    with torch.cuda.stream(st):
        for _ in range(a.warmup):
            out = step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(st)
        for _ in range(a.steps):
            out = step()
        e.record(st)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([s.elapsed_time(e)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        groups, rows_sel = int(out[0].shape[0]), int(out[2].sum().item())
    if rank == 0:
        t = ms.item() / a.steps
        print(json.dumps({"metric": "billion rows/sec on filter + group-by (1e5 int32 keys) + sum/count, 1e9 rows per GPU", "value": n * world / t / 1e6,
                          "unit": "Grows/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": t, "scaling": "weak",
                          "rows_per_gpu": n, "groups": groups, "rows_selected": rows_sel,
                          "merge": "all-gather-v of per-GPU (key, sum, count) rows + re-group on every rank" if world > 1 else "none"}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
