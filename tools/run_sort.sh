set -x
timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sort" 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "binop or division" 2>&1 | tail -3
timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | tee gpurun_out/r02_perf_sort_onesweep.jsonl | tail -5
RFB_SORT_ALGO=lsd timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | tee gpurun_out/r02_perf_sort_lsd.jsonl | tail -5
timeout 600 python tools/perf_ops.py --only timestamp,xbar_time,add_i64 --reps 3 2>&1 | tee gpurun_out/r02_perf_typed.jsonl | tail -6
timeout 600 python -m pytest tests/test_gpu_opslayer.py tests/test_gpu_stats.py -x -q 2>&1 | tail -3
