timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sort" 2>&1 | tail -3
echo "== default (u32 keys 3 CTAs)"; timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
for v in c32_2; do
  echo "== $v"; RFB200_LIB=$PWD/rayforce_b200/librfb200_os_$v.so timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
done
