timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_where_cmp -s 2 -c 1 -f -o gpurun_out/r02_where_cmp python tools/perf_ops.py --reps 1 --only cmp_where_fused > gpurun_out/ncu_w1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_where_mask -s 2 -c 1 -f -o gpurun_out/r02_where_mask python tools/perf_ops.py --reps 1 --only where_mask > gpurun_out/ncu_w2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gather -s 2 -c 1 -f -o gpurun_out/r02_gather python tools/perf_ops.py --reps 1 --only gather_i64 > gpurun_out/ncu_w3.log 2>&1
tail -2 gpurun_out/ncu_w1.log gpurun_out/ncu_w2.log gpurun_out/ncu_w3.log
