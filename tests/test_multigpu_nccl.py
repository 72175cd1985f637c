"""N>1 on hardware: torch.distributed.run with one process per GPU over NCCL (tests/nccl_worker.py).  Needs >= 2 GPUs on
the box (`gpurun --gpus 2`); on a one-GPU box the test is skipped — the host-side sharding / exchange logic is covered on
CPU by tests/test_sharded_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs on one box")
def test_sharded_indexes_over_nccl_match_the_unsharded_oracle():
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and f"nccl worker OK world={world}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
