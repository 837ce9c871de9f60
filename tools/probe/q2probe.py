"""Probe: wall-clock of the multi-key / low-cardinality grouping entry points at 1e7 rows (the H2O Q2 / Q7 building blocks);
used while tracking down per-call cudaMalloc jitter and same-address first-row claims.  python tools/probe/q2probe.py"""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from rayforce_b200 import Context, capi
n=10_000_000
r=np.random.default_rng(1)
torch.cuda.set_device(0)
st=torch.cuda.Stream()
ctx=Context(0, stream=st.cuda_stream)
with torch.cuda.stream(st):
    id1=torch.from_numpy(r.integers(1,101,n)).cuda(); id2=torch.from_numpy(r.integers(1,101,n)).cuda(); v1=torch.from_numpy(r.integers(1,6,n)).cuda()
    ids=[torch.from_numpy(r.integers(1,101,n)).cuda() for _ in range(6)]
st.synchronize()
def t(name, fn, reps=5):
    best=1e9
    with torch.cuda.stream(st):
        for _ in range(reps):
            torch.cuda.synchronize(); t0=time.perf_counter(); out=fn(); ctx.sync(); torch.cuda.synchronize(); best=min(best,(time.perf_counter()-t0)*1e3)
    print(name, round(best,3), "ms wall", flush=True)
    return out
g,f,info=t("group_keys2", lambda: ctx.group_keys([id1,id2]))
t("aggr_sum_1e4", lambda: ctx.aggr(capi.A_SUM, capi.I64, v1, g, info.groups))
t("group_i64 id1", lambda: ctx.group_i64(id1))
fused=(id1*1000+id2)
t("group_i64 fused", lambda: ctx.group_i64(fused))
g7,f7,i7=t("group_keys6", lambda: ctx.group_keys(ids))
print(i7.groups, i7.dense)
