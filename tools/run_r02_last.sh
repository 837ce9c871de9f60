# the driver's round-end sequence on one GPU: smoke, GPU suite, reference arm, our arm (same flags as BENCH_r01.json's cmd)
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_last_smoke.txt 2>&1; tail -4 gpurun_out/r02_last_smoke.txt
( time timeout 900 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02_gpu_suite.txt 2>&1; tail -6 gpurun_out/r02_gpu_suite.txt | head -3
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; tail -4 gpurun_out/r02_bench_reference_arm.err; head -c 300 gpurun_out/r02_bench_reference_arm.json; echo
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -4 gpurun_out/r02_bench_n1.err; head -c 300 gpurun_out/r02_bench_n1.json; echo
