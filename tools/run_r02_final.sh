# final evidence run of round 2 (one B200): operator timings, launch lists, ncu --set full of the new kernels, the reference's own
# harnesses through the drop-in, sanitizer, bench lines.  Outputs land in gpurun_out/ and are summarised into profiles/.
set -x
timeout 1800 python tools/perf_ops.py > gpurun_out/r02_perf_ops.jsonl 2> gpurun_out/r02_perf_ops.err
timeout 600 python tools/perf_ops.py --rows 100000000 --only sort >> gpurun_out/r02_perf_ops.jsonl 2>> gpurun_out/r02_perf_ops.err
tail -3 gpurun_out/r02_perf_ops.err
timeout 900 python tools/h2o_groupby.py > gpurun_out/r02_h2o_groupby.jsonl 2>&1
timeout 900 python tools/join_bench.py > gpurun_out/r02_join_bench.jsonl 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 3 --warmup 3 --no-e2e > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_sort.csv python tools/perf_ops.py --rows 100000000 --only sort_i64_full --reps 1 > gpurun_out/ncu_sl.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_os_pass -s 3 -c 1 -f -o gpurun_out/r02_os_pass python tools/perf_ops.py --rows 100000000 --reps 1 --only sort_i64_full > gpurun_out/ncu_os.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_binop_typed -s 1 -c 1 -f -o gpurun_out/r02_binop_typed python tools/perf_ops.py --reps 1 --only add_timestamp_time > gpurun_out/ncu_bt.log 2>&1
( echo "=== rayforce_tests_dropin"; RFB200_SHIM_STATS=1 timeout 900 oracle/_ref/rayforce_tests_dropin 2>&1 | grep -E "passed|Passed|shim\]" ) > gpurun_out/r02_dropin_reference_tests.txt
( for mode in "" "RFB200_LAZY=1"; do echo "=== rayforce_bench_dropin $mode RFB200_MIN_ROWS=65536"; env $mode RFB200_SHIM_STATS=1 RFB200_MIN_ROWS=65536 timeout 600 oracle/_ref/rayforce_bench_dropin 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "Results|Min Time|Avg Time|shim\] (operator|HBM|lazy)"; done; echo "=== rayforce_bench_ref"; timeout 600 oracle/_ref/rayforce_bench_ref 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "Results|Min Time|Avg Time" ) > gpurun_out/r02_make_bench_1e7.txt
( echo "=== stock"; timeout 900 oracle/_ref/rayforce_ref -f integration/demo/queries.rfl; for mode in "" "RFB200_LAZY=1"; do echo "=== drop-in $mode"; env $mode RFB200_SHIM_STATS=1 timeout 900 oracle/_ref/rayforce_dropin -f integration/demo/queries.rfl; done ) > gpurun_out/r02_dropin_demo_1e8.txt 2>&1
timeout 1200 python tools/sanitizer_workload.py > /dev/null 2>&1 && ( for t in memcheck racecheck synccheck; do echo "=== compute-sanitizer --tool $t"; timeout 1500 compute-sanitizer --tool $t python tools/sanitizer_workload.py 2>&1 | tail -4; done ) > gpurun_out/r02_compute_sanitizer.txt
timeout 1500 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
timeout 1500 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02_gpu_suite.txt 2>&1; tail -4 gpurun_out/r02_gpu_suite.txt
ls -la gpurun_out | tail -24
