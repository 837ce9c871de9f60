#!/usr/bin/env python
"""Per-source-line share of executed warp instructions and of warp-stall samples for one kernel of an ncu report
(ncu --set full --import-source on):   python tools/ncu_lines.py gpurun_out/X.ncu-rep [min_share_pct]"""
import csv
import subprocess
import sys


def num(s):
    try:
        return int(s.replace(",", ""))
    except ValueError:
        return 0


def main():
    path, cut = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    iex, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
    src = [r for r in data if len(r) > max(iex, isamp) and r[0].isdigit()]
    tot, ts = sum(num(r[iex]) for r in src), sum(num(r[isamp]) for r in src)
    print("# %s: %d warp instructions, %d stall samples; lines with >= %.1f%% of either" % (path, tot, ts, cut))
    for r in src:
        ex, sm = num(r[iex]), num(r[isamp])
        if ex >= cut / 100 * tot or sm >= cut / 100 * ts:
            print("%4s  instr %5.1f%%  stalls %5.1f%% | %s" % (r[0], 100 * ex / tot, 100 * sm / ts, r[1].strip()[:120]))


if __name__ == "__main__":
    main()
