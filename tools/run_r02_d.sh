set -x
timeout 2400 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q 2>&1 | tail -40
timeout 900 python -m pytest tests/test_gpu_opslayer.py -m gpu -x -q 2>&1 | tail -15
for mode in "" "RFB200_LAZY=1"; do
echo "=== rayforce_bench_dropin $mode"
env $mode RFB200_SHIM_STATS=1 RFB200_MIN_ROWS=65536 timeout 600 oracle/_ref/rayforce_bench_dropin 2>&1 | grep -E "Results|Min Time|Avg Time|shim\] (operator|HBM|lazy)" 
done
echo "=== rayforce_bench_ref"
timeout 600 oracle/_ref/rayforce_bench_ref 2>&1 | grep -E "Results|Min Time|Avg Time"
