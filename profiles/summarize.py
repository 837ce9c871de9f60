#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files under profiles/.
    python profiles/summarize.py launches gpurun_out/X_launches.csv  > profiles/X_launches.txt
    python profiles/summarize.py full     gpurun_out/X.ncu-rep       > profiles/X_full.txt
"""
import collections
import csv
import subprocess
import sys

FULL_METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
                "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
                "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
                "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
                "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
                "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg = collections.OrderedDict()
    for r in rows:
        ns = float(r[vi].replace(",", ""))
        name = r[ki].split("(")[0][:100]
        bucket = "long(>=500us)" if ns >= 5e5 else "short"
        a = agg.setdefault((name, r[gi], bucket), [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# launches=%d total=%.3f ms" % (len(rows), tot / 1e6))
    for (n, g, b), (c, t) in agg.items():
        print("%5d x %10.1f us avg %6.1f%%  grid=%s %s  %s" % (c, t / c / 1e3, 100 * t / tot, g, b, n))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none; per launch")
    for r in data:
        print("kernel:", r[ki][:160])
        for m in FULL_METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-72s %s %s" % (m, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
