"""The data-parallel training step on real GPUs over NCCL (needs >= 2 GPUs: `gpurun --gpus 2`): gradients of a global
batch split unevenly over the ranks, exchanged by BucketedGradientReducer inside backward(), equal the single-GPU
gradient of the same batch (tests/tools/train_ddp_check.py holds the comparison)."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_bucketed_allreduce_matches_single_gpu_gradient(cuda):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(ROOT, "tests", "tools", "train_ddp_check.py")],
                         capture_output=True, text=True, timeout=600, env=dict(os.environ, DIN_OFFLINE="1"))
    print(out.stdout[-1500:])
    assert out.returncode == 0, out.stderr[-3000:]
    assert "worst relative L2 difference" in out.stdout
