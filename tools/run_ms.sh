for v in default ms_12x2 ms_10x2; do
  echo "== $v"; L=""; [ $v != default ] && L=$PWD/rayforce_b200/librfb200_$v.so
  RFB200_LIB=$L timeout 600 python tools/perf_ops.py --only group_sum_count_i32keys_1e5 --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
done
