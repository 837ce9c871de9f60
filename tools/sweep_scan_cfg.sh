#!/bin/bash
# Tuning sweep for the sum/count scan kernels (run on the GPU box): rebuild k_fold.cu with different launch shapes and
# time the fold / filter+fold operators at 1e9 rows.  Results -> gpurun_out/r01_sweep_scan_cfg.txt
out=gpurun_out/r01_sweep_scan_cfg.txt
: > $out
for cfg in "512 4 4" "256 6 8" "256 8 4" "256 8 2" "1024 2 4" "256 6 4" "384 4 4" "256 4 8"; do
  set -- $cfg
  rm -f rayforce_b200/csrc/build/k_fold.o
  make -s -C rayforce_b200/csrc -j8 CFG="-DSC_THREADS=$1 -DSC_BPS=$2 -DSC_LOADS=$3" all >/dev/null 2>&1 || { echo "cfg $cfg: build failed" >> $out; continue; }
  echo "== threads=$1 ctas_per_sm=$2 loads=$3  $(grep -A2 'k_scan_foldIllLi3ELb1ELb1ELb1E' rayforce_b200/csrc/build/k_fold.ptxas.log | grep -oE '[0-9]+ bytes spill stores' | head -1) $(grep -A3 'k_scan_foldIllLi3ELb1ELb1ELb1E' rayforce_b200/csrc/build/k_fold.ptxas.log | grep -oE 'Used [0-9]+ registers' | head -1)" >> $out
  python tools/perf_ops.py --rows 1000000000 --reps 5 --only sum_i64 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print('   %-28s %8.3f ms %8.1f GB/s' % (d['op'], d['ms_best'], d['GBps']))" >> $out
done
cat $out
