"""Sharded PER on >= 2 GPUs of one node: the peer-mailbox exchange fused into the sample kernel must
give the same indices, weights and rows as the NCCL all-gather path and as the CPU oracle."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_p2p_matches_nccl_and_oracle():
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mp_shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"SHARD_CHECK OK world {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
