#!/bin/bash
# tools/run_plugin_loop.sh N [TAG]: the stock reference binary in plugin mode N times under the fault tracer
# (tools/probe/segv_trace.c, built to tools/probe/segv_trace.so); keeps stdout / stderr of every failing run in
# gpurun_out/plugin_fail_<TAG>_<i>.txt
n=${1:-30}; tag=${2:-a}; fail=0
for i in $(seq 1 $n); do
  LD_PRELOAD=$PWD/tools/probe/segv_trace.so timeout 120 stdbuf -o0 oracle/_ref/rayforce_ref -f integration/demo/plugin.rfl > /tmp/p_$tag.out 2> /tmp/p_$tag.err; rc=$?
  if [ $rc -ne 0 ]; then fail=$((fail+1)); { echo "run $i rc=$rc"; echo "--- stdout"; cat /tmp/p_$tag.out; echo "--- stderr"; cat /tmp/p_$tag.err; } > gpurun_out/plugin_fail_${tag}_$i.txt; fi
done
echo "plugin failures ($tag): $fail / $n"
