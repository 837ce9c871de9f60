timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_stats.py -x -q -k "sort or med or distinct" 2>&1 | tail -3
echo "== in-pass histograms"; timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_sort.csv python tools/perf_ops.py --rows 100000000 --only sort_i64_full --reps 1 > gpurun_out/ncu_sl.log 2>&1
