set -x
ncu --set full --clock-control none --import-source on -k regex:k_where_cmp -s 2 -c 1 -f -o gpurun_out/r01_where_cmp python tools/perf_ops.py --rows 1000000000 --reps 1 --only cmp_where_fused > gpurun_out/ncu_where_cmp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_where_mask -s 2 -c 1 -f -o gpurun_out/r01_where_mask python tools/perf_ops.py --rows 1000000000 --reps 1 --only where_mask > gpurun_out/ncu_where_mask.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fused_accum -s 2 -c 1 -f -o gpurun_out/r01_fused_accum python tools/perf_ops.py --rows 1000000000 --reps 1 --only group_sum_count_i32keys_1e5 > gpurun_out/ncu_fused_accum.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_scatter -s 4 -c 1 -f -o gpurun_out/r01_scatter python tools/perf_ops.py --rows 100000000 --reps 1 --only sort_i64_full > gpurun_out/ncu_scatter.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_fma_fold -s 2 -c 1 -f -o gpurun_out/r01_fma_fold python tools/perf_ops.py --rows 1000000000 --reps 1 --only fma_avg_f64_fused > gpurun_out/ncu_fma.log 2>&1
ls -la gpurun_out/*.ncu-rep
