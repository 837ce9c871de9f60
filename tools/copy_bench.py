#!/usr/bin/env python
"""Host<->device copy rates of the C ABI (rfb_h2d / rfb_d2h) for pageable and pinned host memory."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rayforce_b200 import Context, capi  # noqa: E402

ctx = Context(0)
n = 100_000_000
dev = torch.empty(n, dtype=torch.int64, device="cuda")
for kind in ("pageable", "pinned"):
    h = torch.empty(n, dtype=torch.int64, pin_memory=(kind == "pinned"))
    h.fill_(3)
    out = torch.empty(n, dtype=torch.int64, pin_memory=(kind == "pinned"))
    for rep in range(3):
        t0 = time.perf_counter()
        capi.check(ctx.lib.rfb_h2d(ctx.h, C.c_void_p(dev.data_ptr()), C.c_void_p(h.data_ptr()), n * 8))
        ctx.sync()
        t1 = time.perf_counter()
        capi.check(ctx.lib.rfb_d2h(ctx.h, C.c_void_p(out.data_ptr()), C.c_void_p(dev.data_ptr()), n * 8))
        ctx.sync()
        t2 = time.perf_counter()
        print("%-9s rep %d  h2d %6.1f GB/s   d2h %6.1f GB/s (first rep touches the destination pages)" % (kind, rep, n * 8 / (t1 - t0) / 1e9, n * 8 / (t2 - t1) / 1e9), flush=True)
    assert bool((out == 3).all())
