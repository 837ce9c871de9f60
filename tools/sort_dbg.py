import sys, numpy as np, torch
sys.path.insert(0, '.')
from rayforce_b200 import Context, capi
ctx = Context(0)
r = np.random.default_rng(1)
for t, dt in ((capi.U8, np.uint8), (capi.I16, np.int16), (capi.I32, np.int32), (capi.I64, np.int64)):
    for n in (1, 33, 5000, 100_003):
        col = r.integers(0, 100, n).astype(dt)
        d = torch.from_numpy(col).cuda()
        p = ctx.sort(t, d).cpu().numpy()
        ok = np.array_equal(p, np.argsort(col, kind='stable'))
        print(t, n, ok, flush=True)
