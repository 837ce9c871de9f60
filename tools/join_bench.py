#!/usr/bin/env python
"""The reference's published join benchmarks (docs benchmarks/inner-join.md:29, left-join.md:29: `(ij [id1 id2] x y)` /
`(lj [id1 id2] x y)` on two 1e7-row tables, 1610 ms / 3149 ms, hardware not stated) re-expressed through the C ABI on synthetic
tables of that shape: the row matching (rfb_find_rows_dev / rfb_inner_join_dev) plus the gather of one payload column.
    python tools/join_bench.py [--rows 10000000]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rayforce_b200 import Context, capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    n = a.rows
    r = np.random.default_rng(7)
    torch.cuda.set_device(0)
    st = torch.cuda.Stream()
    ctx = Context(0, stream=st.cuda_stream)
    # x: n rows, keys (id1, id2); y: n rows whose (id1, id2) pairs are unique; ~90 % of x rows have a partner
    y1, y2 = r.permutation(n).astype(np.int64), r.integers(0, 100, n).astype(np.int64)
    pick = r.integers(0, n, n)
    x1, x2 = y1[pick].copy(), y2[pick].copy()
    x2[r.random(n) < 0.1] += 1000
    v = r.uniform(0, 100, n)
    with torch.cuda.stream(st):
        d = {k: torch.from_numpy(t).cuda() for k, t in dict(x1=x1, x2=x2, y1=y1, y2=y2, v=v).items()}
    st.synchronize()

    def left():
        ids = ctx.find_rows([d["y1"], d["y2"]], [d["x1"], d["x2"]])
        return ids

    def inner():
        pi, bi = ctx.inner_join([d["y1"], d["y2"]], [d["x1"], d["x2"]])
        return ctx.gather(capi.F64, d["v"], bi), pi

    for name, fn, pub in (("left_join_row_ids", left, 3149), ("inner_join_rows_plus_payload_gather", inner, 1610)):
        best = 1e9
        with torch.cuda.stream(st):
            for _ in range(a.reps + 1):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(st)
                out = fn()
                e.record(st)
                ctx.sync()
                torch.cuda.synchronize()
                best = min(best, s.elapsed_time(e))
        print(json.dumps({"op": name, "rows_each_side": n, "ms_device_resident": round(best, 3), "reference_published_ms_full_join": pub}), flush=True)
    ids = left().cpu().numpy()
    assert (ids >= 0).sum() > 0.8 * n and np.array_equal(y1[ids[ids >= 0]], x1[ids >= 0])
    ctx.close()


if __name__ == "__main__":
    main()
