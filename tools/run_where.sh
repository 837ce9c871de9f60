timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_opslayer.py tests/test_gpu_fullsize.py -x -q -k "where or compaction or select or filter or fullsize" 2>&1 | tail -3
timeout 300 python tools/perf_ops.py --only where_mask,cmp_where --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
timeout 300 python tools/perf_ops.py --rows 30000000 --only where_mask,cmp_where --reps 5 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
timeout 300 python tools/perf_ops.py --rows 200000000 --only where_mask,cmp_where --reps 5 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
