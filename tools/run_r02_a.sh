set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_group" 2>&1 | tail -8
for v in "" _ms512_4; do
  echo "== variant $v"
  RFB200_LIB=$PWD/rayforce_b200/librfb200$v.so timeout 600 python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 5 2>&1 | tail -2
done
timeout 600 python tools/perf_ops.py --only group_sum_count_i64keys_1e5 --reps 3 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_groupby_v7.csv python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 1 > gpurun_out/ncu_l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ms_scatter -s 2 -c 1 -f -o gpurun_out/r02_ms_scatter_v7 python tools/perf_ops.py --reps 1 --only group_sum_count_i32keys_1e5 > gpurun_out/ncu_ms1.log 2>&1
