set -x
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_mgpu.py -m gpu -x -q 2>&1 | tail -15
for m in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 50 --warmup 5 --no-configs --no-e2e --merge $m 2>&1 | tail -3 | cut -c1-1500
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-configs --no-e2e 2>&1 | tail -1 | cut -c1-600
