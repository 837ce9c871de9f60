timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sort" 2>&1 | tail -3
echo "== default 16x2 staged rids LB4 IL0"; timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
for v in il2 il4 18nr 18nr_il2; do
  echo "== $v"; RFB200_LIB=$PWD/rayforce_b200/librfb200_os_$v.so timeout 300 python tools/perf_ops.py --rows 100000000 --only sort_i64 --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
done
echo "== default again"; timeout 300 python tools/perf_ops.py --rows 100000000 --only sort_i64 --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
