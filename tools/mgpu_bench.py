#!/usr/bin/env python
"""End-to-end timing of the ONE-PROCESS multi-GPU entry points (rfb_mgpu_*: what a host like the reference binds with
rfb_ops_init(host, -1)): a HOST column is cut into row ranges, every device ships its range over its own PCIe link and runs the
fused kernel, the partials are merged on the host.
    python tools/mgpu_bench.py [--rows 1000000000] [--devices 0] [--reps 5] [--pageable]
Prints one JSON line per measurement: `select {(sum x) from t where (< x k)}` (BASELINE config 2) and the group-by of config 4,
host buffers in, host results out, every copy inside the timed region."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (splitmix host fill)
from rayforce_b200 import capi  # noqa: E402
from rayforce_b200.device import MultiGpu  # noqa: E402


def pinned(n, dtype):
    import torch
    t = torch.empty(n, dtype=dtype).pin_memory()
    return t, t.numpy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--devices", type=int, default=0, help="0 = every visible device")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--pageable", action="store_true", help="plain host memory (the library's copier threads stage it)")
    ap.add_argument("--group_rows", type=int, default=250_000_000)
    args = ap.parse_args()
    import torch
    n = args.rows
    mg = MultiGpu(args.devices)
    if args.pageable:
        keep, x = None, np.empty(n, np.int64)
    else:
        keep, x = pinned(n, torch.int64)
    bench.fill_splitmix_host(x, 42, 0, 1 << 40)
    K = 1 << 39
    want_rows, want_sum = 0, 0
    for lo in range(0, n, 1 << 26):                          # the expected answer, in slices (numpy, wrapping sum)
        c = x[lo:lo + (1 << 26)]
        sel = c[c < K]
        want_rows += int(sel.shape[0])
        want_sum = (want_sum + int(sel.sum(dtype=np.int64))) & 0xFFFFFFFFFFFFFFFF
    want_sum = want_sum - (1 << 64) if want_sum >= (1 << 63) else want_sum

    def timed(fn):
        fn()
        ts = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            r = fn()
            ts.append(time.perf_counter() - t0)
        return min(ts), sorted(ts)[len(ts) // 2], r

    best, med, (res, nbytes) = timed(lambda: mg.filter_fold_host(capi.LT, capi.I64, x, K, capi.F_SUM | capi.F_CNT, capi.I64, x))
    assert res.nonnull == want_rows and res.sum == want_sum, (res.nonnull, want_rows, res.sum, want_sum)
    print(json.dumps({"op": "rfb_mgpu_filter_fold_host", "devices": mg.devices, "rows": n, "host_memory": "pageable" if args.pageable else "pinned",
                      "ms_best": round(best * 1e3, 2), "ms_median": round(med * 1e3, 2), "grows_per_s": round(n / best / 1e9, 2),
                      "h2d_gbps": round(nbytes / best / 1e9, 1), "h2d_bytes": nbytes, "result": {"rows": res.nonnull, "sum": res.sum}}), flush=True)
    del x, keep
    # config 4 through the same route: 1e5 int32 keys, i64 values
    m = min(args.group_rows, n)
    if args.pageable:
        kk, k = None, np.empty(m, np.int32)
        kv, v = None, np.empty(m, np.int64)
    else:
        kk, k = pinned(m, torch.int32)
        kv, v = pinned(m, torch.int64)
    k64 = np.empty(m, np.int64)
    bench.fill_splitmix_host(k64, 7, 0, 100_000)
    k[:] = k64
    del k64
    bench.fill_splitmix_host(v, 9, 0, 1 << 20)
    best, med, (gk, gs, gc, nbytes) = timed(lambda: mg.group_sum_count_host(capi.I32, k, v, 100_000))
    assert gk.shape[0] == 100_000 and int(gc.sum()) == m and int(gs.sum()) == int(v.sum())
    print(json.dumps({"op": "rfb_mgpu_group_sum_count_host", "devices": mg.devices, "rows": m, "host_memory": "pageable" if args.pageable else "pinned",
                      "ms_best": round(best * 1e3, 2), "ms_median": round(med * 1e3, 2), "grows_per_s": round(m / best / 1e9, 2),
                      "h2d_gbps": round(nbytes / best / 1e9, 1), "groups": int(gk.shape[0])}), flush=True)
    mg.close()


if __name__ == "__main__":
    main()
