nvidia-smi -L | wc -l
( command -v numactl && numactl -H | head -4 ) 2>&1 | head -6
timeout 600 python tools/mgpu_bench.py --rows 1000000000 --reps 4 2>&1 | tee gpurun_out/r02_mgpu_bench_n8.jsonl | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e 2> gpurun_out/bench8.err | tail -1 > gpurun_out/r02_bench_n8_peer_groups.json
python - <<EOF
import json
d=json.loads(open('gpurun_out/r02_bench_n8_peer_groups.json').read())
print(d['value'], d['ms_per_step'], {k:(v['value'], v['ms_per_step'], v['merge'][:30]) for k,v in d['configs'].items()})
EOF
tail -3 gpurun_out/bench8.err
