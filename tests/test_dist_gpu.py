"""GPU, world_size = 2 over NCCL (needs two GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_dist_gpu.py -m gpu`; skipped on a
one-GPU box): the CUDA kernels + FlatGradBucket.all_reduce() on two query shards reproduce the single-GPU flat gradient of the whole
batch (reference trainer/trainer.py:52-56 + model/loss.py:57: the loss is a SUM over queries, so shard gradients add up)."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs 2 GPUs (NCCL refuses two ranks on one device)")
def test_two_rank_nccl_gradients_equal_the_single_gpu_gradient(tmp_path):
    out = str(tmp_path / "dist_gpu.json")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_gpu_worker.py"), out]
    r = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:]
    res = json.load(open(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dist_gpu_parity.json"), "w") as f:
        json.dump(res, f, indent=1)
    assert set(res) == {"overlap=True", "overlap=False"}
    for k, v in res.items():
        assert v["numel"] == 1755303 and v["world"] == 2
        assert v["max_abs_diff"] <= 2e-6 * v["scale"], (k, v)
        assert v["max_abs_diff_set_to_none"] <= 2e-6 * v["scale"], (k, v)
        assert abs(v["loss_sharded"] - v["loss_single"]) <= 1e-5 * max(1.0, abs(v["loss_single"])), (k, v)
    # with overlap, the segment that completes inside backward is reduced from the side stream, gated on the event the layer-0
    # backward records after its star kernel (both steps of the run), and the reduced gradients are still those of the single GPU
    assert res["overlap=True"]["gated_launches"] >= 1 and res["overlap=False"]["gated_launches"] == 0
