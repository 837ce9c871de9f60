#!/usr/bin/env python
"""First-call vs steady-state wall time of the C-ABI entry points in a fresh process (module loading, workspace allocation):
    python tools/first_call.py [--rows 10000000]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rayforce_b200 import Context, capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    args = ap.parse_args()
    n = args.rows
    t0 = time.perf_counter()
    ctx = Context(0)
    torch.cuda.synchronize()
    print(json.dumps({"op": "context", "ms_first": round((time.perf_counter() - t0) * 1e3, 2)}), flush=True)
    r = np.random.default_rng(1)
    x = torch.from_numpy(r.integers(0, 1 << 40, n).astype(np.int64)).cuda()
    k = torch.from_numpy(r.integers(0, 100_000, n).astype(np.int64)).cuda()
    k32 = k.to(torch.int32)
    f = x.to(torch.float64)
    torch.cuda.synchronize()

    def both(name, fn):
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            fn()
            ctx.sync()
            ts.append((time.perf_counter() - t) * 1e3)
        print(json.dumps({"op": name, "rows": n, "ms_first": round(ts[0], 2), "ms_second": round(ts[1], 2), "ms_third": round(ts[2], 2)}), flush=True)

    both("fold", lambda: ctx.fold(capi.F_SUM | capi.F_CNT, capi.I64, x, n))
    both("filter_fold", lambda: ctx.filter_fold(capi.LT, capi.I64, x, 1 << 39, capi.F_SUM | capi.F_CNT, capi.I64, x, n))
    both("cmp", lambda: ctx.cmp(capi.LT, capi.I64, x, capi.I64, 1 << 39))
    both("cmp_where", lambda: ctx.cmp_where(capi.LT, capi.I64, x, 1 << 39))
    both("binop_add", lambda: ctx.binop(capi.ADD, capi.I64, x, capi.I64, x))
    both("group_i64 (index_group)", lambda: ctx.group_i64(k))
    g, _, info = ctx.group_i64(k)
    both("aggr_sum 1e5 groups", lambda: ctx.aggr(capi.A_SUM, capi.I64, x, g, info.groups))
    both("aggr_count 1e5 groups", lambda: ctx.aggr(capi.A_COUNT, capi.I64, x, g, info.groups))
    both("group_sum_count i32 keys", lambda: ctx.group_sum_count(capi.I32, k32, x, 100_000))
    both("group_sum_count i64 keys", lambda: ctx.group_sum_count(capi.I64, k, x, 100_000))
    both("sort i64", lambda: ctx.sort(capi.I64, x))
    both("fma_fold", lambda: ctx.fma_fold(capi.F_SUM | capi.F_CNT, f, f, f, n))
    ctx.close()


if __name__ == "__main__":
    main()
