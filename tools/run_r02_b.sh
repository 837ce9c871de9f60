set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_group" 2>&1 | tail -3
for v in "" _v7 _a512x2s3 _a1024x2s2 _a512x2s2; do
  echo "== variant $v"
  RFB200_LIB=$PWD/rayforce_b200/librfb200$v.so timeout 600 python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 5 2>&1 | tail -2
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_groupby_v8.csv python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 1 > gpurun_out/ncu_l.log 2>&1
for v in _a1024x2s2 _a512x2s3; do
RFB200_LIB=$PWD/rayforce_b200/librfb200$v.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_groupby_v8$v.csv python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 1 > gpurun_out/ncu_l.log 2>&1
done
timeout 1200 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -c 3000 gpurun_out/bench_r02_a.json; tail -5 gpurun_out/bench_r02_a.err
