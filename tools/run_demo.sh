# the drop-in demonstration at 1e8 rows on a fresh box: stock vs drop-in vs drop-in with lazy results (profiles/r02_dropin_demo_1e8.txt)
( echo "# '(3 runs)' lines are the TOTAL of three evaluations"; echo "=== stock"; timeout 900 oracle/_ref/rayforce_ref -f integration/demo/queries.rfl; for mode in "" "RFB200_LAZY=1"; do echo "=== drop-in $mode"; env $mode RFB200_SHIM_STATS=1 timeout 900 oracle/_ref/rayforce_dropin -f integration/demo/queries.rfl; done ) > gpurun_out/r02_dropin_demo_1e8.txt 2>&1
sed 's/\x1b\[[0-9;]*m//g' gpurun_out/r02_dropin_demo_1e8.txt | grep -E "===|runs\)|run\)"
