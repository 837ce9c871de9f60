timeout 300 python -m pytest tests/test_gpu_ops.py -x -q -k "group_merge_over_peer" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_nccl.py tests/test_gpu_mgpu.py -x -q 2>&1 | tail -8
for m in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e --merge $m 2> gpurun_out/bench2_$m.err | tail -1 > gpurun_out/r02_bench_n2_$m.json
python - <<EOF
import json
d=json.loads(open('gpurun_out/r02_bench_n2_$m.json').read())
print('$m', d['ms_per_step'], {k:(v['ms_per_step'], v['merge'][:40]) for k,v in d['configs'].items()})
EOF
done
tail -3 gpurun_out/bench2_peer.err
