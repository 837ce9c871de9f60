#!/usr/bin/env python
"""bench.py — headline benchmark: 1e9-row int64 filter (x < k) + sum per B200 (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--rows R] [--impl ours|reference]

A "step" is one pass of the hot path over the GPU's row-range shard of the column:
    select {s: (sum x) from: t where: (< x k)}        (reference: core/cmp.c -> core/ops.c ops_where ->
                                                        core/rayforce.c at_ids -> core/math.c ray_sum)
executed as ONE fused sm_100a kernel (rfb_filter_fold_dev), its 3-number result read back to the host, and for N > 1
one NCCL all-reduce of the per-GPU (rows, sum) partials.  Prints ONE JSON line (rank 0):

  value      billion rows/s over all GPUs, column resident in HBM when the timed region starts (CUDA events, max over ranks)
  roofline   the fused kernel alone: algorithmic bytes (8 B/row) / mean CUDA-event duration of the launches inside the
             timed region, against MEASURED_PEAKS.json hbm_gbs
  e2e        same metric through the host-pointer C-ABI call (rfb_filter_fold_host): pinned HOST column in, chunked
             cudaMemcpyAsync + kernels inside the timed region, host result out
  cpu_baseline  the reference's own CPU implementation (oracle/_ref, compiled from the reference sources) running the same
             Rayfall select on a bounded sample, all host cores, on rank 0 at N = 1

`--impl reference` times only that CPU arm (rank 0) and prints the same line shape with "impl": "reference".
Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 42
MODULUS = 1 << 40          # x[i] = splitmix64(SEED, global_row) mod 2^40
K_CONST = 1 << 39          # predicate x < 2^39 -> 50 % selectivity
GOLDEN = 0x9E3779B97F4A7C15
METRIC = "billion rows/sec on 1e9-row int64 filter+sum"
UNIT = "Grows/s"
CPU_SAMPLE_ROWS = 100_000_000
COLLECT_EVERY = 10         # device-resident loop: the host waits for (and checks) the result of every 10th pass; the passes in
                           # between are queued behind each other on the stream, each still delivering its merged result to host memory


def shifted_seed(seed: int, first_row: int) -> int:
    """splitmix64(seed', i) == splitmix64(seed, first_row + i)"""
    return (seed + first_row * GOLDEN) & 0xFFFFFFFFFFFFFFFF


def splitmix_column(seed: int, first_row: int, n: int, modulus: int) -> np.ndarray:
    """numpy restatement of the synthetic column (same generator as rfb_fill_splitmix_dev / rfo_splitmix64)"""
    out = np.empty(n, np.int64)
    step = 1 << 24
    with np.errstate(over="ignore"):
        for lo in range(0, n, step):
            hi = min(n, lo + step)
            i = np.arange(first_row + lo + 1, first_row + hi + 1, dtype=np.uint64)
            z = np.uint64(seed) + i * np.uint64(GOLDEN)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            out[lo:hi] = (z % np.uint64(modulus)).astype(np.int64)
    return out


# ---------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    """samples SM clock + throttle reasons of one GPU while the timed regions run (NVML, ~2 ms period)"""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                 0x80: "hw_power_brake_slowdown"}
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------- CPU reference arm

def mem_available_gb() -> float:
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_sample_rows(full_rows: int, runs: int) -> int:
    """The CPU arm runs the FULL row count when the host has the memory for the reference's intermediates (8 B column + 1 B
    mask + 8 B ids and 8 B gathered values per selected row: ~17 GB per 1e9 rows at 50 % selectivity) and the run
    stays within a few minutes (~4 s per 1e9 rows on 16 cores); otherwise a prefix of 1e8 rows, scaled by the metric's unit."""
    need_gb = full_rows * 30e-9 + 8
    if mem_available_gb() >= need_gb and runs * full_rows * 4.5e-9 <= 240:
        return full_rows
    return min(full_rows, CPU_SAMPLE_ROWS)


def fill_splitmix_host(dst: np.ndarray, seed: int, first_row: int, modulus: int):
    """dst[i] = splitmix64(seed, first_row + i) mod modulus, in place, numpy in slices on a few threads"""
    n = dst.shape[0]
    step = 1 << 24

    def work(lo):
        hi = min(n, lo + step)
        dst[lo:hi] = splitmix_column(seed, first_row + lo, hi - lo, modulus)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(work, range(0, n, step)))


def cpu_reference_run(steps: int, warmup: int, sample_rows: int = CPU_SAMPLE_ROWS):
    """Times the reference's own CPU path: the Rayfall query (select {s: (sum x) from: t where: (< x k)}) through the
    reference evaluator compiled from its sources (oracle/_ref/librayforce_ref.so, all host cores).  Falls back to the
    single-threaded C port (oracle/) when the compiled reference is not present.
    -> dict(value, seconds_per_step, kind, cores, sample, check)"""
    from oracle import bindings as ob
    times = []
    if ob.Reference.available():
        R = ob.Reference.get()
        o = R.eval("(set x (til %d))" % sample_rows)
        dst = np.frombuffer((C.c_char * (sample_rows * 8)).from_address(o + 16), dtype=np.int64)
        fill_splitmix_host(dst, SEED, 0, MODULUS)
        expect_sum = 0
        for lo in range(0, sample_rows, 1 << 26):
            c = dst[lo:lo + (1 << 26)]
            expect_sum = (expect_sum + int(c[c < K_CONST].sum(dtype=np.int64))) & 0xFFFFFFFFFFFFFFFF
        if expect_sum >= 1 << 63:
            expect_sum -= 1 << 64
        R.eval("(set t (table [x] (list x)))")
        q = "(select {s: (sum x) from: t where: (< x %d)})" % K_CONST
        got = None
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            r = R.eval(q)
            dt = time.perf_counter() - t0
            cols = R.list_items(R.list_items(r)[1])
            got = int(R.to_numpy(cols[0], drop=False)[0][0])
            R.drop(r)
            if i >= warmup:
                times.append(dt)
        kind, cores = "reference", R.cores
        del dst
        R.eval("(set t 0)")
        R.eval("(set x 0)")
    else:
        sample_rows = min(sample_rows, CPU_SAMPLE_ROWS)
        col = splitmix_column(SEED, 0, sample_rows, MODULUS)
        expect_sum = int(col[col < K_CONST].sum(dtype=np.int64))
        O = ob.Oracle()
        got = None
        for i in range(max(1, warmup // 3) + max(1, steps // 3)):
            t0 = time.perf_counter()
            m = O.cmp(ob.LT, ob.I64, col, ob.I64, K_CONST)
            ids = O.where(m)
            g = O.at_ids(ob.I64, col, ids)
            got = int(O.fold(ob.SUM, ob.I64, g)[0])
            dt = time.perf_counter() - t0
            if i >= max(1, warmup // 3):
                times.append(dt)
        kind, cores = "port", 1
    if got != expect_sum:
        raise SystemExit("CPU reference arm produced %d, expected %d" % (got, expect_sum))
    sec = statistics.mean(times)
    what = "the FULL %d-row column" % sample_rows if sample_rows >= 1_000_000_000 else "a %d-row prefix of the same splitmix64 column (scaled by rows/s)" % sample_rows
    return {"value": sample_rows / sec / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": "%s, Rayfall select through the reference evaluator, mean of %d runs (best %.1f ms); host MemAvailable %.0f GB"
                      % (what, len(times), min(times) * 1e3, mem_available_gb()),
            "seconds_per_step": sec, "rows": sample_rows}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(args.steps, args.warmup, cpu_sample_rows(args.rows, args.steps + args.warmup))
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def ncu_traffic_bytes(rows):
    """DRAM bytes (read + write) per launch of the dominant kernel from the committed `ncu --set full` capture
    (profiles/*scan_fold*_full.txt, taken at 1e9 rows); None when no capture matches this row count"""
    import glob
    import re
    if rows != 1_000_000_000:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*scan_fold*_full.txt")), reverse=True):
        txt = open(path).read()
        rd = re.search(r"dram__bytes_read\.sum\s+([0-9.]+)\s+(\w+)", txt)
        wr = re.search(r"dram__bytes_write\.sum\s+([0-9.]+)\s+(\w+)", txt)
        if rd and wr:
            return float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)], os.path.relpath(path, ROOT)
    return None, None


MERGE_TEXT = {"peer": "one 32-thread kernel per step over NVLink peer mailboxes (rfb_fold_allreduce_peers: P2P stores + sequence flags, "
                      "merged (nonnull, sum) written straight into mapped host memory); NCCL carries the group-by merges of configs 4-5",
              "nccl": "one NCCL all-reduce of (nonnull, sum) per step + device-to-host copy"}


def workload_config(args, world, merge_mode="nccl"):
    return {"workload": "select {s: (sum x) from: t where: (< x k)}: int64 column, x = splitmix64(42, row) mod 2^40, "
                        "k = 2^39 (50%% selectivity), %d rows per GPU sharded by row range" % args.rows,
            "rows_per_gpu": args.rows, "global_rows": args.rows * world, "selectivity": 0.5,
            "l2_policy": "inputs (8 GB per GPU) far exceed the 126 MB L2; no flush needed",
            "host_collects": "every pass writes its (merged) result into mapped host memory from the GPU; the host waits for and "
                             "checks every %dth pass and the last one, the passes in between are queued on the stream" % COLLECT_EVERY,
            "merge": "none (1 GPU)" if world == 1 else MERGE_TEXT.get(merge_mode, merge_mode)}



# ---------------------------------------------------------------------------------------------- BASELINE configs 3, 4, 5

def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


E2E_CONFIG_ROWS = 250_000_000


def run_extra_configs(ctx, stream, dev, n, rank, world, K, W, with_e2e=True, merge_mode="nccl"):
    """BASELINE.json configs[2..4] on the same box, same process, device-resident, each timed with CUDA events around K steps
    (max over ranks): fp64 (a*b+c) -> avg; group-by 1e5 int32 keys sum/count; filter + group-by + sum sharded by row range with
    the NCCL merge.  A step of the group-by configs is the whole rfb_group_sum_count_dev call (sample, scatter, accumulate,
    first rows, emit, its host synchronisations) — `call_ms`, not one kernel.  -> dict for the JSON line's "configs" key."""
    import torch
    import torch.distributed as dist
    from rayforce_b200 import capi, shard
    peak, _ = peak_hbm()
    out = {}
    first_row = rank * n

    def timed(step):
        for _ in range(max(W, 3)):
            res = step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = ctx.launches
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream)
        for _ in range(K):
            res = step()
        e.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = s.elapsed_time(e) / K
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, (ctx.launches - l0) // K

    def entry(workload, ms, alg_bytes_per_gpu, launches, result, merge):
        gbs = alg_bytes_per_gpu / (ms * 1e-3) / 1e9
        return {"workload": workload, "value": n * world / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": K,
                "rows_per_gpu": n, "n_gpus": world, "gpu_launches_per_step": int(launches),
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "algorithmic_bytes_per_step_per_gpu": alg_bytes_per_gpu, "timed": "whole call (all its kernels and host syncs)"},
                "merge": merge, "result": result}

    ne = min(n, E2E_CONFIG_ROWS)      # rows of the end-to-end leg of configs 3-5 (pinned host columns: 24 B/row for config 3)

    def timed_e2e(step, Ke=2):
        """host columns in, host result out, copies inside: wall clock around Ke calls, max over ranks"""
        res = step()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(Ke):
            res = step()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - w0) * 1e3 / Ke
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    def e2e_entry(ms, h2d, d2h, api):
        return {"value": ne * world / (ms * 1e-3) / 1e9, "unit": UNIT, "rows_per_gpu": ne, "ms_per_step": ms, "steps": 2,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "api": api}

    def pinned_copy(t):
        h = torch.empty(ne, dtype=t.dtype, pin_memory=True)
        with torch.cuda.stream(stream):
            h.copy_(t[:ne], non_blocking=True)
        stream.synchronize()
        return h

    # ---- config 3: (avg (+ (* a b) c)) over three F64 columns, one fused kernel (24 B/row)
    with torch.cuda.stream(stream):
        a, b, c = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    for col, seed in ((a, 1), (b, 2), (c, 3)):
        ctx.fill_splitmix(capi.F64, col, n, shifted_seed(seed, first_row), 1 << 20, 0, 0, float(1 << 20))
    ctx.sync()

    def step_fma():
        if world > 1 and merge_mode == "peer":               # error-free (hi, lo) partial sums folded in rank order by one tiny kernel
            ctx.fma_fold_async(capi.F_SUM | capi.F_CNT, a, b, c, n)
            m = ctx.fold_allreduce_peers(capi.F64)
            return m.sum / m.nonnull
        r = ctx.fma_fold(capi.F_SUM | capi.F_CNT, a, b, c, n)
        if world == 1:
            return r.sum / r.nonnull
        tot = shard.allgather_sum_f64(r.sum, dev)            # fp64 partials added in rank order: independent of the reduction tree
        cnt = torch.tensor([r.nonnull], dtype=torch.int64, device=dev)
        dist.all_reduce(cnt)
        return tot / int(cnt.item())
    ms, avg, launches = timed(step_fma)
    assert 0.74 < avg < 0.76, avg                             # E[a*b + c] = 1/4 + 1/2 for uniform [0, 1) columns
    out["fma_avg"] = entry("(avg (+ (* a b) c)): three F64 columns, splitmix64 / 2^20 in [0, 1), fused k_fma_fold", ms, 24 * n, launches,
                           {"avg": avg}, "none" if world == 1 else ("peer mailboxes: (hi, lo) partial sums folded in rank order by one 32-thread kernel" if merge_mode == "peer"
                                                                   else "all-gather of (sum, count) partials, added in rank order"))
    if with_e2e:
        ha, hb, hc = (pinned_copy(t) for t in (a, b, c))
        na, nb_, nc = ha.numpy(), hb.numpy(), hc.numpy()

        def step_fma_host():
            r, nbytes = ctx.fma_fold_host(capi.F_SUM | capi.F_CNT, na, nb_, nc)
            return r.sum / r.nonnull, nbytes
        ems, (eavg, nbytes) = timed_e2e(step_fma_host)
        assert 0.74 < eavg < 0.76, eavg
        out["fma_avg"]["e2e"] = e2e_entry(ems, nbytes, 72, "rfb_fma_fold_host (three pinned host columns -> cudaMemcpyAsync -> fused kernel -> host result); first %d rows per GPU" % ne)
        del ha, hb, hc, na, nb_, nc
    del a, b, c

    # ---- config 4 / 5: group-by 1e5 int32 keys, sum + count of an i64 column; config 5 adds the filter and the multi-GPU merge
    with torch.cuda.stream(stream):
        k = torch.empty(n, dtype=torch.int32, device=dev)
        v = torch.empty(n, dtype=torch.int64, device=dev)
    ctx.fill_splitmix(capi.I32, k, n, shifted_seed(7, first_row), 100_000, 0, 0)
    ctx.fill_splitmix(capi.I64, v, n, shifted_seed(9, first_row), 1 << 20, 0, 0)
    ctx.sync()
    regroup = shard.gpu_regroup(ctx)
    # N > 1: the exchange step of the group-by.  Default (--merge peer): every rank's (key, sum, count) lists are read in place over
    # NVLink peer memory by the merge kernels (rfb_group_merge_peers); --merge nccl, or no CUDA IPC: all-gather + re-group.
    group_peer = False
    if world > 1 and merge_mode == "peer":
        try:
            ctx.peer_groups_setup(rank, world, 1 << 18)
            group_peer = True
        except Exception as e:
            print("[bench] group exchange buffers unavailable (%s); NCCL all-gather instead" % e, file=sys.stderr)
        flag = torch.tensor([1 if group_peer else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # all ranks or none
        group_peer = int(flag.item()) == 1

    def grouped(filtered):
        def step():
            if filtered:
                lk, ls, lc = ctx.group_sum_count(capi.I32, k, v, 100_000, capi.LT, capi.I64, v, 1 << 19)
            else:
                lk, ls, lc = ctx.group_sum_count(capi.I32, k, v, 100_000)
            if world > 1:
                with torch.cuda.stream(stream):
                    if group_peer:
                        return shard.merge_group_partials_peers(ctx, lk, ls, lc, 1 << 18, regroup)
                    return shard.merge_group_partials(lk, ls, lc, regroup)
            return lk, ls, lc
        return step
    merge = "none" if world == 1 else ("every GPU's (key, sum, count) rows read in place over NVLink peer memory by the merge kernels (rfb_group_merge_peers): "
                                       "direct-address fold + compaction in global first-occurrence order on every rank, no collective" if group_peer
                                       else "one all-gather of every GPU's (key, sum, count) rows + re-group on every rank (NCCL)")
    with torch.cuda.stream(stream):
        ms, (gk, gs, gc), launches = timed(grouped(False))
        groups, rows, total = int(gk.shape[0]), int(gc.sum().item()), int(gs.sum().item())
    assert groups == 100_000 and rows == n * world, (groups, rows)
    out["groupby_1e5"] = entry("select {s: (sum v) c: (count v) from t by k}: 1e5 int32 keys (splitmix64 mod 1e5), i64 values < 2^20", ms, 12 * n,
                               launches, {"groups": groups, "rows": rows, "sum_of_sums": total}, merge)
    with torch.cuda.stream(stream):
        ms, (gk, gs, gc), launches = timed(grouped(True))
        groups, rows, total = int(gk.shape[0]), int(gc.sum().item()), int(gs.sum().item())
    assert groups == 100_000 and 0.49 * n * world < rows < 0.51 * n * world, (groups, rows)
    out["filter_groupby_sharded"] = entry("select {s: (sum v) c: (count v) from t by k where (< v 2^19)}: rows sharded by row range over the GPUs", ms,
                                          12 * n, launches, {"groups": groups, "rows_selected": rows, "sum_of_sums": total}, merge)
    if with_e2e:
        hk, hv = pinned_copy(k), pinned_copy(v)
        nk, nv = hk.numpy(), hv.numpy()
        for name, filtered in (("groupby_1e5", False), ("filter_groupby_sharded", True)):
            def step_group_host():
                if filtered:
                    gk_, gs_, gc_, nbytes = ctx.group_sum_count_host(capi.I32, nk, nv, 100_000, capi.LT, capi.I64, nv, 1 << 19)
                else:
                    gk_, gs_, gc_, nbytes = ctx.group_sum_count_host(capi.I32, nk, nv, 100_000)
                return int(gk_.shape[0]), int(gc_.sum()), nbytes
            ems, (eg, erows, nbytes) = timed_e2e(step_group_host)
            assert eg == 100_000 and (0.49 * ne < erows < 0.51 * ne if filtered else erows == ne), (eg, erows)
            out[name]["e2e"] = e2e_entry(ems, nbytes, 3 * 8 * eg, "rfb_group_sum_count_host (pinned host key + value columns -> cudaMemcpyAsync -> fused group-by -> "
                                         "host group lists; per-GPU lists, no cross-GPU merge in this leg); first %d rows per GPU" % ne)
        del hk, hv, nk, nv
    del k, v
    return out


# ---------------------------------------------------------------------------------------------- GPU arm

def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from rayforce_b200 import Context, capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local, stream=stream.cuda_stream)
    n = args.rows
    K, W = args.steps, args.warmup

    # --- the shard: rows [rank*n, (rank+1)*n) of the global column, generated straight into HBM
    with torch.cuda.stream(stream):
        x = torch.empty(n, dtype=torch.int64, device=dev)
        res = torch.zeros(16, dtype=torch.int64, device=dev)     # rfb_fold_t image (72 bytes) for the device-side merge
    res_host = torch.zeros(2, dtype=torch.int64).pin_memory()    # merged (nonnull, sum) lands here
    ctx.fill_splitmix(capi.I64, x, n, shifted_seed(SEED, rank * n), MODULUS)
    ctx.sync()

    # N > 1: the final merge of the per-GPU (count, sum) partials.  Default: one 32-thread kernel over NVLink peer mailboxes
    # (rfb_fold_allreduce_peers: P2P stores + sequence flags, merged result straight into mapped host memory).  --merge nccl,
    # or a box without CUDA IPC, takes the NCCL all-reduce + device-to-host copy instead.
    merge_mode = "none (1 GPU)"
    if world > 1:
        merge_mode = "nccl"
        if args.merge == "peer":
            try:
                ctx.peer_mailbox_setup(rank, world)
                merge_mode = "peer"
            except Exception as e:                                 # noqa: BLE001
                if rank == 0:
                    print("peer mailboxes unavailable (%s): NCCL all-reduce" % e, file=sys.stderr)
            flag = torch.tensor([1 if merge_mode == "peer" else 0], dtype=torch.int64, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)            # all ranks or none
            merge_mode = "peer" if int(flag.item()) == 1 else "nccl"

    def step_resident(ev_s=None, ev_e=None, collect=True):
        """one pass, column resident in HBM -> (rows, sum) on the host.  Every pass ends with the GPU writing its (merged)
        result into mapped host memory; collect=False only skips the HOST's wait for it, so that the next pass is already
        queued behind this one (a stream of queries) — the host then reads the result of a later pass (COLLECT_EVERY)."""
        if world > 1 and merge_mode == "peer":
            if ev_s is not None:
                ev_s.record(stream)
            ctx.filter_fold_async(capi.LT, capi.I64, x, K_CONST, capi.F_SUM | capi.F_CNT, capi.I64, x, n)
            if ev_e is not None:
                ev_e.record(stream)
            ctx.fold_allreduce_peers_async(capi.I64)
            if not collect:
                return None
            r = ctx.fold_peers_result(capi.I64)
            return r.nonnull, r.sum
        if world == 1:
            if ev_s is not None:
                ev_s.record(stream)
            ctx.filter_fold_async(capi.LT, capi.I64, x, K_CONST, capi.F_SUM | capi.F_CNT, capi.I64, x, n)
            if ev_e is not None:
                ev_e.record(stream)
            if not collect:
                return None
            r = ctx.fold_result(capi.I64)
            return r.nonnull, r.sum
        ctx.set_result_ptr(res)
        if ev_s is not None:
            ev_s.record(stream)
        ctx.filter_fold_async(capi.LT, capi.I64, x, K_CONST, capi.F_SUM | capi.F_CNT, capi.I64, x, n)
        if ev_e is not None:
            ev_e.record(stream)
        with torch.cuda.stream(stream):
            dist.all_reduce(res[1:3])                             # nonnull, sum_i64 (wraps mod 2^64)
            res_host.copy_(res[1:3], non_blocking=True)
        stream.synchronize()
        ctx.set_result_ptr(None)
        return int(res_host[0]), int(res_host[1])

    sampler = ClockSampler(local)

    # --- device-resident timing
    for _ in range(max(W, 3)):
        first = step_resident()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = ctx.launches
    if rank == 0:
        sampler.start()
    t_start.record(stream)
    for i in range(K):
        r = step_resident(*evs[i], collect=(i % COLLECT_EVERY == COLLECT_EVERY - 1 or i == K - 1))
        if r is not None:
            assert r == first, "non-deterministic result %r vs %r at step %d" % (r, first, i)
            got = r
    t_end.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = ctx.launches - launches0
    total_ms = t_start.elapsed_time(t_end)
    kernel_ms = statistics.mean(s.elapsed_time(e) for s, e in evs)
    assert got == first, "non-deterministic result %r vs %r" % (got, first)

    # --- end to end: pinned HOST column in, host result out, copies inside the timed region
    e2e = None
    if not args.no_e2e:
        hx_t = torch.empty(n, dtype=torch.int64, pin_memory=True)
        with torch.cuda.stream(stream):
            hx_t.copy_(x, non_blocking=True)
        stream.synchronize()
        hx = hx_t.numpy()
        Ke = max(1, min(K, args.e2e_steps))

        def step_e2e():
            r, nbytes = ctx.filter_fold_host(capi.LT, capi.I64, hx, K_CONST, capi.F_SUM | capi.F_CNT, capi.I64, hx)
            rows, s = r.nonnull, r.sum
            if world > 1:
                t = torch.tensor([rows, s], dtype=torch.int64).to(dev)
                dist.all_reduce(t)
                rows, s = (int(v) for v in t.cpu())
            return (rows, s), nbytes

        for _ in range(2):
            ge, nbytes = step_e2e()
        assert ge == first, "host-layer result %r differs from device-layer result %r" % (ge, first)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        le0 = ctx.launches
        e0.record(stream)
        w0 = time.perf_counter()
        for _ in range(Ke):
            step_e2e()
        e1.record(stream)
        torch.cuda.synchronize()
        w1 = time.perf_counter()
        e2e_ms = max(e0.elapsed_time(e1), (w1 - w0) * 1e3)       # the copies run on a second stream: take the larger
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e = {"value": n * world * Ke / (e2e_ms * 1e-3) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
               "d2h_bytes_per_step": 72 * ((n * 8 + (64 << 20) - 1) // (64 << 20)), "steps": Ke,
               "ms_per_step": e2e_ms / Ke, "launches_per_step": (ctx.launches - le0) // Ke,
               "api": "rfb_filter_fold_host (pinned host column -> chunked cudaMemcpyAsync + fused kernel -> host result)"}
        del hx, hx_t
    configs = None
    if not args.no_configs:
        del x
        configs = run_extra_configs(ctx, stream, dev, n, rank, world, max(1, min(K, args.config_steps)), W, with_e2e=not args.no_e2e, merge_mode=merge_mode)
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, kernel_ms = (float(v) for v in t.cpu())

    if rank == 0:
        peaks, peak_src = None, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
            peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            peak = 6650.0
        achieved = 8.0 * n / (kernel_ms * 1e-3) / 1e9
        traffic, traffic_src = (args.ncu_traffic, "--ncu-traffic") if args.ncu_traffic else ncu_traffic_bytes(n)
        line = {"metric": METRIC, "value": n * world * K / (total_ms * 1e-3) / 1e9, "unit": UNIT, "n_gpus": world,
                "steps": K, "warmup": max(W, 3), "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": workload_config(args, world, merge_mode),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": traffic, "traffic_source": traffic_src, "kernel": "k_scan_fold<i64,i64,SUM|CNT,pred,same-column,null-free predicate>",
                             "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": 8 * n, "peak_source": peak_src},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "configs": configs,
                "result": {"rows_selected_nonnull": int(first[0]), "sum": int(first[1])}}
        if world == 1 and not args.no_cpu:
            try:
                r = cpu_reference_run(steps=3, warmup=1, sample_rows=cpu_sample_rows(n, 4))
                line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the checker is optional at bench time; say so rather than hide it
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--merge", choices=["peer", "nccl"], default="peer", help="N > 1: how the per-GPU (count, sum) partials are merged each step")
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 3-5 (fma->avg, group-by, sharded filter+group-by)")
    ap.add_argument("--config-steps", type=int, default=10, help="timed steps of each of the configs 3-5")
    ap.add_argument("--ncu-traffic", type=float, default=None,
                    help="dram bytes per launch of the dominant kernel from the committed ncu capture (profiles/)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
