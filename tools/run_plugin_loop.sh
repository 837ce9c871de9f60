fail=0
for i in $(seq 1 25); do
  timeout 120 oracle/_ref/rayforce_ref -f integration/demo/plugin.rfl > /tmp/p.out 2> /tmp/p.err; rc=$?
  if [ $rc -ne 0 ]; then fail=$((fail+1)); echo "run $i rc=$rc"; tail -3 /tmp/p.out; tail -5 /tmp/p.err; fi
done
echo "failures: $fail / 25"
which gdb catchsegv 2>/dev/null
