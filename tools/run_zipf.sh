set -x
timeout 600 python tools/perf_ops.py --only zipf --reps 3 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_zipf.csv python tools/perf_ops.py --only zipf --reps 1 > gpurun_out/ncu_z.log 2>&1
