"""Multi-GPU data plane on hardware: tests/nccl_worker.py under torchrun on 2 GPUs (NCCL over NVLink).
Skipped on a single-GPU box; run with ``gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu``."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_nccl_data_plane_matches_single_gpu(gpu_pkg):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (found %d)" % n)
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("NCCL_WORKER_OK") == world, r.stdout[-3000:]
