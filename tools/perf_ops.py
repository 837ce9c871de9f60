#!/usr/bin/env python
"""Per-operator device timings (CUDA events, column resident in HBM) for DESIGN.md's kernel table.
    python tools/perf_ops.py [--rows 1000000000] [--reps 5]
Prints one JSON object per operator: ms, algorithmic bytes, achieved GB/s and fraction of the measured HBM peak."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayforce_b200 import Context, capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    n = args.rows
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    ctx = Context(0, stream=st.cuda_stream)
    dev = "cuda:0"

    def col(t, seed, modulus, offset=0, scale=1.0, null_every=0):
        x = torch.empty(n, dtype={capi.I64: torch.int64, capi.I32: torch.int32, capi.F64: torch.float64}[t], device=dev)
        ctx.fill_splitmix(t, x, n, seed, modulus, offset, null_every, scale)
        return x

    def timed(name, fn, alg_bytes, note=""):
        if args.only and not any(o in name for o in args.only.split(",")):
            return
        for _ in range(2):
            fn()
        ctx.sync()
        ms = []
        for _ in range(args.reps):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(st)
            fn()
            e.record(st)
            ctx.sync()
            torch.cuda.synchronize()
            ms.append(s.elapsed_time(e))
        best = min(ms)
        gbs = alg_bytes / (best * 1e-3) / 1e9
        print(json.dumps({"op": name, "rows": n, "ms_best": round(best, 3), "ms_median": round(sorted(ms)[len(ms) // 2], 3),
                          "grows_per_s": round(n / best / 1e6, 2), "alg_bytes": alg_bytes, "GBps": round(gbs, 1),
                          "frac_of_measured_hbm": round(gbs / peak, 3), "note": note}), flush=True)

    with torch.cuda.stream(st):
        x = col(capi.I64, 42, 1 << 40)
        K = 1 << 39
        timed("sum_i64", lambda: ctx.fold(capi.F_SUM | capi.F_CNT, capi.I64, x, n), 8 * n)
        timed("minmax_i64", lambda: ctx.fold(capi.F_MIN | capi.F_MAX, capi.I64, x, n), 8 * n)
        timed("filter_sum_i64_same_col", lambda: ctx.filter_fold(capi.LT, capi.I64, x, K, capi.F_SUM | capi.F_CNT, capi.I64, x, n), 8 * n)
        y = col(capi.I64, 43, 1 << 20)
        timed("filter_sum_i64_two_cols", lambda: ctx.filter_fold(capi.LT, capi.I64, x, K, capi.F_SUM | capi.F_CNT, capi.I64, y, n), 16 * n)
        timed("and_filter_2cols_sum", lambda: ctx.multi_filter_fold([(capi.LT, capi.I64, x, K), (capi.GE, capi.I64, y, 1 << 19)], True,
                                                                   capi.F_SUM | capi.F_CNT, capi.I64, y, n), 16 * n,
              "where (and (< x k1) (>= y k2)), sum y: fused, 16 B/row")
        timed("cmp_lt_mask", lambda: ctx.cmp(capi.LT, capi.I64, x, capi.I64, K), 9 * n, "8 B in + 1 B mask out")
        mask = ctx.cmp(capi.LT, capi.I64, x, capi.I64, K)
        sel = int(ctx.where(mask).shape[0])
        timed("where_mask", lambda: ctx.where(mask), n + 8 * sel, "1 B in + 8 B per selected row out")
        timed("cmp_where_fused", lambda: ctx.cmp_where(capi.LT, capi.I64, x, K), 8 * n + 8 * sel)
        ids = ctx.where(mask)
        del mask
        timed("gather_i64", lambda: ctx.gather(capi.I64, x, ids), 24 * sel, "ids 8 + col 8 + out 8 per selected row")
        timed("gather_fold_i64", lambda: ctx.gather_fold(capi.F_SUM | capi.F_CNT, capi.I64, x, ids, sel), 16 * sel)
        del ids
        timed("add_i64_vv", lambda: ctx.binop(capi.ADD, capi.I64, x, capi.I64, y), 24 * n)
        timed("mul_i64_va", lambda: ctx.binop(capi.MUL, capi.I64, x, capi.I64, 3), 16 * n)
        timed("div_i64_va", lambda: ctx.binop(capi.DIV, capi.I64, x, capi.I64, 7), 16 * n)
        # the typed matrix kernels (k_binop_typed): the I64 / I32 payloads reinterpreted as TIMESTAMP / TIME columns
        tcol = col(capi.I32, 44, 86_400_000)
        timed("add_timestamp_time_vv", lambda: ctx.binop(capi.ADD, capi.TIMESTAMP, x, capi.TIME, tcol), 20 * n, "timestamp + time (ms -> ns): 8 + 4 in, 8 out")
        timed("sub_timestamp_timestamp_vv", lambda: ctx.binop(capi.SUB, capi.TIMESTAMP, x, capi.TIMESTAMP, y), 24 * n)
        timed("xbar_timestamp_va", lambda: ctx.binop(capi.XBAR, capi.TIMESTAMP, x, capi.I64, 60_000_000_000), 16 * n, "generic 64-bit division inside the typed kernel")
        timed("xbar_time_va", lambda: ctx.binop(capi.XBAR, capi.TIME, tcol, capi.I32, 60_000), 8 * n)
        timed("xbar_time_i64atom", lambda: ctx.binop(capi.XBAR, capi.TIME, tcol, capi.I64, 60_000), 8 * n, "`xbar time 60000`: the literal is an i64 atom")
        del tcol
        del x
        # config 3: fp64 a*b+c -> avg
        a, b, c = col(capi.F64, 1, 1 << 20, 0, float(1 << 20)), col(capi.F64, 2, 1 << 20, 0, float(1 << 20)), col(capi.F64, 3, 1 << 20, 0, float(1 << 20))
        timed("fma_avg_f64_fused", lambda: ctx.fma_fold(capi.F_SUM | capi.F_CNT, a, b, c, n), 24 * n, "config 3 fused")

        def unfused():
            t1, _ = ctx.binop(capi.MUL, capi.F64, a, capi.F64, b)
            t2, _ = ctx.binop(capi.ADD, capi.F64, t1, capi.F64, c)
            return ctx.fold(capi.F_SUM | capi.F_CNT, capi.F64, t2, n)
        timed("fma_avg_f64_operator_at_a_time", unfused, 24 * n, "same work as three operators (64 B/row of traffic)")
        timed("floor_f64", lambda: ctx.unop_f64(capi.FLOOR, a), 16 * n)
        del a, b, c
        # config 4: group-by 1e5 keys sum + count
        k32 = col(capi.I32, 7, 100_000)
        timed("group_sum_count_i32keys_1e5", lambda: ctx.group_sum_count(capi.I32, k32, y, 100_000), 12 * n, "config 4 fused (2 passes over keys)")
        timed("group_sum_count_i32keys_1e5_where", lambda: ctx.group_sum_count(capi.I32, k32, y, 100_000, capi.LT, capi.I64, y, 1 << 19), 12 * n)
        del k32
        # config 4's contention variant: Zipf-like keys, P(k) ~ 1 / k over [0, 1e5) (k = floor(1e5^u) - 1, u uniform)
        if not args.only or "zipf" in args.only.split(","):
            kz = torch.empty(n, dtype=torch.int32, device=dev)
            step = 1 << 26
            for i0 in range(0, n, step):
                m = min(step, n - i0)
                u = torch.rand(m, device=dev, dtype=torch.float64)
                kz[i0:i0 + m] = (torch.pow(100_000.0, u).to(torch.int64) - 1).clamp_(0, 99_999).to(torch.int32)
                del u
            torch.cuda.synchronize()
            timed("group_sum_count_i32keys_1e5_zipf", lambda: ctx.group_sum_count(capi.I32, kz, y, 100_000), 12 * n, "config 4, Zipf(1) keys: hot keys contend")
            del kz
        k64 = col(capi.I64, 7, 100_000)
        timed("group_sum_count_i64keys_1e5", lambda: ctx.group_sum_count(capi.I64, k64, y, 100_000), 16 * n)
        timed("index_group_i64_dense_1e5", lambda: ctx.group_i64(k64), 16 * n, "scope + claim + number + assign (group_ids written)")
        gids, firsts, info = ctx.group_i64(k64)
        timed("aggr_sum_i64_1e5", lambda: ctx.aggr(capi.A_SUM, capi.I64, y, gids, info.groups), 16 * n)
        timed("aggr_avg_i64_1e5", lambda: ctx.aggr(capi.A_AVG, capi.I64, y, gids, info.groups), 16 * n)
        del gids, firsts, k64
        # sparse key domains (range >> 2^28): shared-memory open-addressing tables + device-wide merge (k_hash_group.cu)
        for card in (1000, 100_000, 10_000_000):
            kd = col(capi.I64, 13, card)
            ks, _ = ctx.binop(capi.MUL, capi.I64, kd, capi.I64, 0x9E3779B97F4A7C1)      # distinct multiples of a large odd constant
            del kd
            timed("group_sum_count_sparse_i64keys_%d" % card, lambda: ctx.group_sum_count(capi.I64, ks, y, card), 16 * n,
                  "sparse keys: per-CTA smem hash tables (TMA-staged tiles) + device-wide merge")
            del ks
        k100 = col(capi.I64, 9, 100)
        timed("group_sum_count_i64keys_100", lambda: ctx.group_sum_count(capi.I64, k100, y, 100), 16 * n, "low cardinality (H2O id1-like)")
        for card in (1, 2, 8):
            kc = col(capi.I64, 9, card)
            timed("group_sum_count_i64keys_%d" % card, lambda: ctx.group_sum_count(capi.I64, kc, y, card), 16 * n, "very low cardinality")
            del kc
        g100, _, i100 = ctx.group_i64(k100)
        timed("aggr_sum_i64_100", lambda: ctx.aggr(capi.A_SUM, capi.I64, y, g100, i100.groups), 16 * n, "CTA-private 32-bit shared accumulators")
        timed("aggr_avg_i64_100", lambda: ctx.aggr(capi.A_AVG, capi.I64, y, g100, i100.groups), 16 * n)
        timed("aggr_dev_i64_100", lambda: ctx.aggr(capi.A_DEV, capi.I64, y, g100, i100.groups), 16 * n, "f64 L2 atomics")
        del k100, g100
    if n <= 250_000_000 or "sort" in args.only.split(","):
        with torch.cuda.stream(st):
            ks = col(capi.I64, 11, 0)
            timed("sort_i64_full_width", lambda: ctx.sort(capi.I64, ks), 8 * n, "8 radix passes")
            k40 = col(capi.I64, 11, 1 << 32)
            timed("sort_i64_32bit_range", lambda: ctx.sort(capi.I64, k40), 8 * n, "4 passes (constant digits skipped)")
            del k40
            kf = col(capi.F64, 12, 1 << 40, 0, 1024.0)
            timed("sort_f64", lambda: ctx.sort(capi.F64, kf), 8 * n)
            k32 = col(capi.I32, 13, 0)
            timed("sort_i32", lambda: ctx.sort(capi.I32, k32), 4 * n, "4 passes")
    ctx.close()


if __name__ == "__main__":
    main()
