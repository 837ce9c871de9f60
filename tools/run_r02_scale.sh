# N-GPU evidence run (gpurun --gpus N): NCCL / mgpu parity, then the driver's own scaling command at N GPUs, both merges
set -x
N=${1:-8}
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_mgpu.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 1500 gpurun_out/r02_bench_n$N.json; tail -3 gpurun_out/r02_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 50 --warmup 5 --no-configs --no-e2e --merge nccl > gpurun_out/r02_bench_n${N}_nccl.json 2> gpurun_out/r02_bench_n${N}_nccl.err
head -c 300 gpurun_out/r02_bench_n${N}_nccl.json
