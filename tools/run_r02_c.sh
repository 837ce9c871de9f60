set -x
timeout 1200 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_group or aggr" 2>&1 | tail -8
timeout 600 python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 5 2>&1 | tail -2
timeout 600 python tools/perf_ops.py --only aggr_ --reps 3 2>&1 | tail -8
timeout 600 python tools/perf_ops.py --only group_sum_count_i64keys_1e5 --reps 3 2>&1 | tail -1
