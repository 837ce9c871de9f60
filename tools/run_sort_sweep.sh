timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_stats.py tests/test_gpu_opslayer.py -x -q -k "sort or med or distinct or asc or desc" 2>&1 | tail -3
timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
