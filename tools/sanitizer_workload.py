#!/usr/bin/env python
"""One small invocation of every kernel family (for compute-sanitizer memcheck / racecheck / synccheck runs)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rayforce_b200 import Context, capi  # noqa: E402

ctx = Context(0)
r = np.random.default_rng(1)
n = 70_003


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


x = dev(r.integers(-1000, 1000, n).astype(np.int64))
y = dev(r.integers(-1000, 1000, n).astype(np.int64))
f = dev(r.uniform(-1, 1, n))
k = dev(r.integers(0, 500, n).astype(np.int64))
kw = dev(r.integers(-(1 << 60), 1 << 60, n).astype(np.int64))
k32 = dev(r.integers(0, 5000, n).astype(np.int32))
ctx.fold(capi.F_ALL, capi.I64, x, n)
ctx.fold(capi.F_SUM | capi.F_CNT, capi.F64, f, n)
ctx.filter_fold(capi.LT, capi.I64, x, 10, capi.F_SUM | capi.F_CNT, capi.I64, x, n)
ctx.filter_fold(capi.GE, capi.I64, x, 10, capi.F_ALL, capi.F64, f, n)
ctx.fma_fold(capi.F_ALL, f, f, f, n)
m = ctx.cmp(capi.LT, capi.I64, x, capi.I64, y)
ids = ctx.where(m)
ctx.mask_logic(capi.M_AND, m, m)
ids2 = ctx.cmp_where(capi.GT, capi.I64, x, 0)
ctx.gather(capi.I64, x, ids)
ctx.gather_fold(capi.F_ALL, capi.I64, x, ids2, ids2.shape[0])
ctx.binop(capi.DIV, capi.I64, x, capi.I64, y)
ctx.binop(capi.FDIV, capi.F64, f, capi.I64, 3)
ctx.unop_f64(capi.FLOOR, f)
for keys in (k, kw):
    g, fi, info = ctx.group_i64(keys)
    for op in (capi.A_SUM, capi.A_MIN, capi.A_MAX, capi.A_COUNT, capi.A_AVG):
        ctx.aggr(op, capi.I64, x, g, info.groups)
    ctx.aggr(capi.A_SUM, capi.F64, f, g, info.groups)
ctx.group_keys([k, k])
ctx.group_sum_count(capi.I32, k32, x, 5000)
ctx.group_sum_count(capi.I64, k, x, 500, capi.LT, capi.I64, x, 100)
# fused group-by: every accumulate strategy (shared memory / scope-free residue / key-range partitions with the on-demand block
# store / L2 atomics), with and without a predicate, aligned and unaligned columns
kbig = dev(r.integers(-3000, 40_000, n + 1).astype(np.int64))
for strat in ("smem", "part", "l2", None):
    if strat:
        os.environ["RFB_GROUP_STRATEGY"] = strat
    else:
        os.environ.pop("RFB_GROUP_STRATEGY", None)
        os.environ["RFB_PART_MIN_ROWS"] = "1000"
    if strat != "smem":
        ctx.group_sum_count(capi.I64, kbig[:n], x, 50_000)
        ctx.group_sum_count(capi.I64, kbig[1:], x, 50_000, capi.LT, capi.I64, x, 100)
    ctx.group_sum_count(capi.I32, k32, x, 5000)
    ctx.group_sum_count(capi.I64, k, x, 500, capi.GE, capi.I64, x, -100)
os.environ.pop("RFB_PART_MIN_ROWS", None)
# multi-key row hashing, med / dev / row lists, fp64 moments, joins
wide = [dev(r.integers(0, 7, n).astype(np.int64) << 50), dev(r.integers(0, 9, n).astype(np.int64) << 45)]
ctx.group_keys(wide)
g, fi, info = ctx.group_i64(k)
ctx.aggr(capi.A_MED, capi.I64, x, g, info.groups)
ctx.aggr(capi.A_MED, capi.F64, f, g, info.groups)
ctx.aggr(capi.A_DEV, capi.I64, x, g, info.groups)
ctx.aggr(capi.A_AVG, capi.F64, f, g, info.groups)
g100 = dev(r.integers(0, 100, n).astype(np.int64))
ctx.aggr(capi.A_DEV, capi.F64, f, g100, 100)
ctx.aggr(capi.A_SUM, capi.F64, f, g100, 100)
ctx.aggr(capi.A_AVG, capi.I64, x, g100, 100)
ctx.group_rows(g, info.groups)
ctx.med(capi.I64, x)
ctx.stddev(capi.I64, x)
ctx.stddev(capi.F64, f)
ctx.find_rows([k, y], [k, x])
ctx.inner_join([kw], [kw])
# asof / window joins (right side ordered by key, time), distinct
order = np.lexsort((r.integers(0, 10_000, n), k.cpu().numpy()))
rk, rtime = dev(k.cpu().numpy()[order]), dev(np.sort(r.integers(0, 10_000, n)).astype(np.int64))
ctx.asof_join([rk], capi.I64, rtime, [k], dev(r.integers(0, 10_000, n).astype(np.int64)))
rt32 = dev(np.sort(r.integers(0, 10_000, n)).astype(np.int32))
wlo = dev(r.integers(0, 9_000, n).astype(np.int32))
for op in (capi.A_SUM, capi.A_MIN, capi.A_COUNT, capi.A_AVG):
    ctx.window_join(op, capi.I64, x, [dev(np.sort(k.cpu().numpy()))], rt32, [k], wlo, wlo, 0)
ctx.distinct(k)
ctx.sort(capi.I64, x)
ctx.sort(capi.F64, f, True)
h = r.integers(-1000, 1000, 2_000_000).astype(np.int64)
ctx.filter_fold_host(capi.LT, capi.I64, h, 10, capi.F_ALL, capi.I64, h, chunk_rows=300_000)
# round 2: narrow (<= 32-partition, ballot-ranked, TMA-staged) group-by passes with packed 32 / 64-bit records and the exception
# list, sparse-key hash path, grouped sums through the partition passes, aggr_first / aggr_last, constant-divisor division
os.environ["RFB_PART_MIN_ROWS"] = "1000"
nn = 300_007
kn = dev(r.integers(0, 100_000, nn).astype(np.int32))
vn = r.integers(0, 1 << 20, nn).astype(np.int64)
vn[::997] = capi.NULL_I64
vn = dev(vn)
ctx.group_sum_count(capi.I32, kn, vn, 100_008)
ctx.group_sum_count(capi.I32, kn, vn, 100_008, capi.LT, capi.I64, vn, 1 << 19)
vw = dev(r.integers(-(1 << 44), 1 << 44, nn).astype(np.int64))
ctx.group_sum_count(capi.I64, dev(r.integers(-70_000, 130_000, nn).astype(np.int64)), vw, 200_008)
os.environ["RFB_GROUP_STRATEGY"] = "hash"
ctx.group_sum_count(capi.I64, dev((r.integers(0, 5000, nn).astype(np.int64)) * 0x9E3779B97F4A7C1), vn, 5008)
os.environ.pop("RFB_GROUP_STRATEGY", None)
gn = dev(r.integers(0, 60_000, nn).astype(np.int64))
ctx.aggr(capi.A_SUM, capi.I64, vn, gn, 60_000)
ctx.aggr(capi.A_AVG, capi.I64, vw, gn, 60_000)
os.environ.pop("RFB_PART_MIN_ROWS", None)
ctx.aggr(capi.A_FIRST, capi.I64, x, g, info.groups)
ctx.aggr_last(capi.F64, f, g, info.groups, 8)
for op in (capi.DIV, capi.MOD, capi.XBAR):
    ctx.binop(op, capi.I64, x, capi.I64, -7)
# round 2, later: single-sweep sort passes (TMA-staged tiles, decoupled look-back) for 64- and 32-bit key words, both directions;
# the histogram + scatter passes they replace; the typed arithmetic kernels (temporal units, narrow integers, B8)
for col, t in ((kw, capi.I64), (k32, capi.I32), (f, capi.F64), (dev(r.integers(0, 200, n).astype(np.uint8)), capi.U8)):
    ctx.sort(t, col)
    ctx.sort(t, col, True)
os.environ["RFB_SORT_ALGO"] = "lsd"
ctx.sort(capi.I64, kw)
os.environ.pop("RFB_SORT_ALGO", None)
ts = dev(r.integers(0, 1 << 50, n).astype(np.int64))
tm = dev(r.integers(0, 86_400_000, n).astype(np.int32))
dt = dev(r.integers(0, 20_000, n).astype(np.int32))
i16 = dev(r.integers(-300, 300, n).astype(np.int16))
u8 = dev(r.integers(0, 256, n).astype(np.uint8))
ctx.binop(capi.ADD, capi.TIMESTAMP, ts, capi.TIME, tm)
ctx.binop(capi.ADD, capi.DATE, dt, capi.TIME, tm)
ctx.binop(capi.SUB, capi.TIMESTAMP, ts, capi.TIMESTAMP, ts)
ctx.binop(capi.XBAR, capi.TIMESTAMP, ts, capi.I64, 60_000_000_000)
ctx.binop(capi.XBAR, capi.TIME, tm, capi.I32, 60_000)
ctx.binop(capi.MUL, capi.I16, i16, capi.I16, i16)
ctx.binop(capi.DIV, capi.U8, u8, capi.U8, u8)
ctx.binop(capi.ADD, capi.U8, u8, capi.F64, 2.5)
ctx.binop(capi.MOD, capi.I16, i16, capi.I64, 7)
# the group-by exchange over peer memory with a world of one (publish / meet / fold / emit), 8-byte keys through the narrow passes
import ctypes as C  # noqa: E402
hnd = (C.c_char * 64)()
capi.check(ctx.lib.rfb_peer_groups_create(ctx.h, 100_000, hnd))
capi.check(ctx.lib.rfb_peer_groups_bind(ctx.h, 0, 1, C.c_char_p(bytes(hnd.raw))))
mk = dev(r.integers(-50, 20_000, 60_000).astype(np.int64))
ctx.group_merge_peers(mk, x[:60_000].contiguous(), y[:60_000].contiguous(), 30_000)
ctx.group_merge_peers(mk, x[:60_000].contiguous(), y[:60_000].contiguous(), 30_000)
os.environ["RFB_PART_MIN_ROWS"] = "1000"
ctx.group_sum_count(capi.I64, dev(r.integers(0, 100_000, nn).astype(np.int64)), vn, 100_008)
os.environ.pop("RFB_PART_MIN_ROWS", None)
ctx.sync()
print("sanitizer workload done, launches:", ctx.launches)
