set -x
timeout 900 python -m pytest tests/test_gpu_ops.py -x -q -k "binop or division" 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_opslayer.py -x -q 2>&1 | tail -5
timeout 600 python tools/perf_ops.py --only timestamp,xbar_time,add_i64,zipf --reps 3 2>&1 | tee gpurun_out/r02_perf_typed.jsonl | tail -8
