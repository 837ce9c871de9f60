set -x
RFB200_LAZY=1 RFB200_LAZY_MIN=16384 RFB200_SHIM_STATS=1 timeout 900 oracle/_ref/rayforce_tests_dropin 2>&1 | sed 's/\x1b\[[0-9;]*m//g' | grep -E "passed|Passed|Failed|FAIL|lazy|operator calls|HBM" | tail -12
oracle/_ref/rayforce_ref -f integration/demo/parity.rfl > /tmp/p0.txt 2>&1
RFB200_LAZY=1 RFB200_LAZY_MIN=16384 RFB200_SHIM_STATS=1 oracle/_ref/rayforce_dropin -f integration/demo/parity.rfl > /tmp/p1.txt 2> /tmp/p1.err
grep " : " /tmp/p0.txt > /tmp/a.txt; grep " : " /tmp/p1.txt > /tmp/b.txt; diff /tmp/a.txt /tmp/b.txt && echo PARITY_IDENTICAL_WITH_LAZY; grep -E "lazy|HBM|operator calls" /tmp/p1.err; tail -3 /tmp/p1.txt
