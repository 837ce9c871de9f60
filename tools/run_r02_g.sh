set -x
timeout 2400 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_opslayer.py tests/test_gpu_join.py -m gpu -q 2>&1 | tail -25
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "division or binop" 2>&1 | tail -3
timeout 600 python tools/perf_ops.py --only div_ --reps 3 2>&1 | tail -4
RFB200_SHIM_STATS=1 oracle/_ref/rayforce_dropin -f integration/demo/parity.rfl 2>&1 | grep -E "window|launched|ray_min|ray_max|ray_avg|ray_sum" | head
