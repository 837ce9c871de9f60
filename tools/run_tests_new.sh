timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "sort" 2>&1 | tail -2
