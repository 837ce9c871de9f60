timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "where or compaction" 2>&1 | tail -5
echo "== stream"; timeout 300 python tools/perf_ops.py --only where_mask,cmp_where --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
echo "== tile"; RFB_WHERE_ALGO=tile timeout 300 python tools/perf_ops.py --only where_mask,cmp_where --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
