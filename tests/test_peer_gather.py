"""Multi-GPU scan with the gather fused into the kernel (records stored into rank 0's array over NVLink peer memory): needs two
GPUs on the box, skipped otherwise.  The host-side sharding logic itself is covered on CPU by test_distributed_gloo.py."""
import os
import subprocess
import sys

import pytest


@pytest.mark.gpu
def test_peer_gather_equals_nccl_gather():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29577", os.path.join(root, "tests", "peer_gather_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0 and "PEER_GATHER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
