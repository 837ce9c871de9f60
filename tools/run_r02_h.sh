set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_group or aggr" 2>&1 | tail -3
timeout 600 python tools/perf_ops.py --only group_sum_count_i --reps 5 2>&1 | tail -3 | cut -c1-130
timeout 600 python tools/perf_ops.py --only aggr_sum_i64_1e5 --reps 3 2>&1 | tail -1 | cut -c1-130
timeout 1200 python tools/sanitizer_workload.py > /dev/null 2>&1 && ( for t in memcheck racecheck synccheck; do echo "=== compute-sanitizer --tool $t"; timeout 1500 compute-sanitizer --tool $t python tools/sanitizer_workload.py 2>&1 | tail -4; done ) > gpurun_out/r02_compute_sanitizer.txt
cat gpurun_out/r02_compute_sanitizer.txt
