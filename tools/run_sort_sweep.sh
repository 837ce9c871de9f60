timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_stats.py tests/test_gpu_opslayer.py -x -q -k "sort or med or distinct or asc or desc" 2>&1 | tail -2
for v in default os32_24x2r1 os32_28x2r0 os32_32x2r0; do
  echo "== $v"; L=""; [ $v != default ] && L=$PWD/rayforce_b200/librfb200_$v.so
  RFB200_LIB=$L timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
done
