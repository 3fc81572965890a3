"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): the in-kernel NVLink all-reduce of sb_closure_peer /
sb_fit_step against the NCCL path and the single-GPU result, bitwise equality across ranks, graph replay, uneven shards
(tools/peer_multi_check.py under torchrun, 2 ranks). The combine logic itself is covered on CPU by test_dist_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_exchange_two_ranks():
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "peer_multi_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "PEER TEST PASSED" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
