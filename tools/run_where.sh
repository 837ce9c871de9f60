timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_opslayer.py -x -q -k "where or compaction or select or filter" 2>&1 | tail -2
for r in 20000000 60000000 100000000; do
timeout 300 python tools/perf_ops.py --rows $r --only where_mask,cmp_where --reps 5 2>&1 | grep -o '"rows": [0-9]*\|"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
RFB200_LIB=$PWD/rayforce_b200/librfb200_prev.so timeout 300 python tools/perf_ops.py --rows $r --only where_mask,cmp_where --reps 5 2>&1 | grep -o '"rows": [0-9]*\|"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
done
