"""The one collective of the path over NCCL: needs >= 2 GPUs (skipped otherwise; run with
`gpurun --gpus 2 -- python -m pytest tests/test_nccl_gpu.py -m gpu`).  The CPU twin (gloo, world size 2) is
tests/test_host_logic.py::test_gather_poses_world2_gloo."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_gather_poses_over_nccl_returns_every_ranks_poses_in_image_order(cuda_dev):
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 4 if n_gpus >= 4 else 2
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
                          os.path.join(root, "tests", "nccl_gather_check.py")], capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, (out.stdout[-2000:], out.stderr[-3000:])
    assert "NCCL_GATHER_OK world=%d" % world in out.stdout
