timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "binop or division" 2>&1 | tail -3
timeout 600 python tools/perf_ops.py --only xbar_time --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*\|"frac_of_measured_hbm": [0-9.]*' | tr '\n' ' '; echo
