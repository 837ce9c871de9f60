fail=0
for i in $(seq 1 40); do timeout 120 oracle/_ref/rayforce_ref -f integration/demo/plugin.rfl > /tmp/p.out 2> /tmp/p.err; rc=$?; if [ $rc -ne 0 ]; then fail=$((fail+1)); echo "run $i rc=$rc"; fi; done
echo "plugin failures: $fail / 40"
for i in 1 2; do ( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02_gpu_suite.txt 2>&1; tail -5 gpurun_out/r02_gpu_suite.txt | head -2; done
