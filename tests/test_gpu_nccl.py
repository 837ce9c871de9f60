"""Multi-process NCCL parity: tools/multi_gpu_check.py under torchrun on every GPU of the box (2 .. 8) — rank r holds
rows [r*n, (r+1)*n); the merged filter+fold (one all-reduce) and filter+group-by+sum (one all-gather + re-group) results must
equal what ONE GPU computes over all the rows, bit for bit, including the global first-occurrence group order.  Skipped on a
1-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu` runs it."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_results_equal_the_single_gpu_results_over_nccl():
    import torch
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--rows", "5000000"]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=900)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0, out[-2000:] + r.stderr.decode(errors="replace")[-3000:]
    rep = json.loads([l for l in out.splitlines() if l.startswith("{")][-1])
    assert rep["ok"] and rep["world"] == world and rep["group_rows_equal"] and rep["groups"] == 100_000
    assert rep["merged_fold"] == rep["single_gpu_fold"]
    assert rep["peer_mailbox_allreduce_equal"] is True, rep        # the NVLink mailbox merge gives the NCCL merge's bits
    assert rep["peer_group_merge_equal"] is True, rep              # so does the group-by merge over peer memory (rfb_group_merge_peers)
