timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "sort" 2>&1 | tail -3
echo "== default 16x2 staged rids LB4"; timeout 300 python tools/perf_ops.py --rows 100000000 --only sort --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
for v in lb8 lb1 18nr 16nr; do
  echo "== $v"; RFB200_LIB=$PWD/rayforce_b200/librfb200_os_$v.so timeout 300 python tools/perf_ops.py --rows 100000000 --only sort_i64 --reps 3 2>&1 | grep -o '"op": "[a-z0-9_]*"\|"ms_best": [0-9.]*' | tr '\n' ' '; echo
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_os_pass -s 3 -c 1 -f -o gpurun_out/r02_os_pass python tools/perf_ops.py --rows 100000000 --reps 1 --only sort_i64_full > gpurun_out/ncu_os.log 2>&1
