timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_opslayer.py -x -q -k "group or aggr or narrow or part" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 600 python tools/perf_ops.py --only group_sum_count_i32keys_1e5,group_sum_count_i64keys_1e5,aggr_sum_i64_1e5,aggr_avg_i64_1e5 --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
