#!/usr/bin/env python
"""The reference's PUBLISHED benchmark for this path: the 7 H2O db-benchmark group-by queries on G1_1e7_1e2_0_0
(reference docs/docs/content/get-started/benchmarks/group-by.md:32,54-60; hardware not stated there), re-expressed
through the C ABI on synthetic data of the same shape (1e7 rows, K = 100; symbols as interned i64 ids).

    python tools/h2o_groupby.py [--rows 10000000] [--reps 5]

For every query: device-resident time (CUDA events, columns in HBM) and end-to-end time (the query's columns copied from
pinned host memory first), next to the reference's published milliseconds.  Results are cross-checked with numpy."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayforce_b200 import Context, capi  # noqa: E402

PUBLISHED_MS = {"Q1": 60, "Q2": 74, "Q3": 118, "Q4": 72, "Q5": 122, "Q6": 104, "Q7": 1394}
QUERY = {"Q1": "sum v1 by id1", "Q2": "sum v1 by id1,id2", "Q3": "sum v1, avg v3 by id3", "Q4": "avg v1,v2,v3 by id4",
         "Q5": "sum v1,v2,v3 by id6", "Q6": "max v1 - min v2 by id3", "Q7": "sum v3, count by id1..id6"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    n, K = args.rows, 100
    r = np.random.default_rng(108)
    host = {"id1": r.integers(1, K + 1, n), "id2": r.integers(1, K + 1, n), "id3": r.integers(1, n // K + 1, n),
            "id4": r.integers(1, K + 1, n), "id5": r.integers(1, K + 1, n), "id6": r.integers(1, n // K + 1, n),
            "v1": r.integers(1, 6, n), "v2": r.integers(1, 16, n)}
    host = {k: v.astype(np.int64) for k, v in host.items()}
    host["v3"] = np.round(r.uniform(0, 100, n), 6)
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    ctx = Context(0, stream=st.cuda_stream)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
    with torch.cuda.stream(st):
        d = {k: v.to("cuda", non_blocking=True) for k, v in pinned.items()}
    st.synchronize()
    I64, F64 = capi.I64, capi.F64

    def q1(c):
        return ctx.group_sum_count(I64, c["id1"], c["v1"], K + 1)

    def q2(c):
        g, f, info = ctx.group_keys([c["id1"], c["id2"]])
        return ctx.aggr(capi.A_SUM, I64, c["v1"], g, info.groups)[0], f

    def q3(c):
        g, f, info = ctx.group_i64(c["id3"])
        return ctx.aggr(capi.A_SUM, I64, c["v1"], g, info.groups)[0], ctx.aggr(capi.A_AVG, F64, c["v3"], g, info.groups)[0], f

    def q4(c):
        g, f, info = ctx.group_i64(c["id4"])
        return [ctx.aggr(capi.A_AVG, t, c[v], g, info.groups)[0] for v, t in (("v1", I64), ("v2", I64), ("v3", F64))], f

    def q5(c):
        g, f, info = ctx.group_i64(c["id6"])
        return [ctx.aggr(capi.A_SUM, t, c[v], g, info.groups)[0] for v, t in (("v1", I64), ("v2", I64), ("v3", F64))], f

    def q6(c):
        g, f, info = ctx.group_i64(c["id3"])
        mx, mn = ctx.aggr(capi.A_MAX, I64, c["v1"], g, info.groups)[0], ctx.aggr(capi.A_MIN, I64, c["v2"], g, info.groups)[0]
        return ctx.binop(capi.SUB, I64, mx, I64, mn)[0], f

    def q7(c):
        g, f, info = ctx.group_keys([c["id%d" % i] for i in range(1, 7)])
        return ctx.aggr(capi.A_SUM, F64, c["v3"], g, info.groups)[0], ctx.aggr(capi.A_COUNT, F64, c["v3"], g, info.groups)[0], f

    qs = {"Q1": (q1, ["id1", "v1"]), "Q2": (q2, ["id1", "id2", "v1"]), "Q3": (q3, ["id3", "v1", "v3"]),
          "Q4": (q4, ["id4", "v1", "v2", "v3"]), "Q5": (q5, ["id6", "v1", "v2", "v3"]), "Q6": (q6, ["id3", "v1", "v2"]),
          "Q7": (q7, ["id1", "id2", "id3", "id4", "id5", "id6", "v3"])}

    # --- correctness spot checks against numpy
    k, s, c = q1(d)
    want = np.bincount(host["id1"], weights=host["v1"]).astype(np.int64)
    assert np.array_equal(s.cpu().numpy(), want[k.cpu().numpy()]), "Q1 mismatch"
    s2, f2 = q2(d)
    fk = host["id1"] * 1000 + host["id2"]
    uk, inv = np.unique(fk, return_inverse=True)
    w2 = np.bincount(inv, weights=host["v1"]).astype(np.int64)
    got_keys = fk[f2.cpu().numpy()]
    assert np.array_equal(s2.cpu().numpy(), w2[np.searchsorted(uk, got_keys)]), "Q2 mismatch"
    s7, c7, f7 = q7(d)
    assert int(c7.sum().item()) == n and abs(float(s7.sum().item()) - float(host["v3"].sum())) < 1e-3 * n, "Q7 mismatch"

    with torch.cuda.stream(st):
        for name, (fn, cols) in qs.items():
            best_dev, best_e2e = 1e9, 1e9
            for _ in range(args.reps + 1):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(st)
                fn(d)
                e_.record(st)
                ctx.sync()
                torch.cuda.synchronize()
                best_dev = min(best_dev, s_.elapsed_time(e_))
            for _ in range(args.reps):
                s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s_.record(st)
                fresh = {c: pinned[c].to("cuda", non_blocking=True) for c in cols}
                fn(fresh)
                e_.record(st)
                ctx.sync()
                torch.cuda.synchronize()
                best_e2e = min(best_e2e, s_.elapsed_time(e_))
            print(json.dumps({"query": name, "rayfall": QUERY[name], "rows": n, "ms_device_resident": round(best_dev, 3),
                              "ms_end_to_end_from_pinned_host": round(best_e2e, 3), "reference_published_ms": PUBLISHED_MS[name],
                              "published_over_e2e": round(PUBLISHED_MS[name] / best_e2e, 1)}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
