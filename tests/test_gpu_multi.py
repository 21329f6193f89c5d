"""Sharded PER on >= 2 GPUs of one node: the peer-mailbox exchanges (statistics published by the mutating
kernels / by the sample kernel) must give the same indices, weights and rows as the NCCL all-gather path and
as the CPU oracle, and an unmatched call must time out instead of hanging (tests/mp_shard_worker.py)."""
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
    world = int(os.environ.get("MRL_TEST_WORLD", 0)) or (2 if n < 4 else (4 if n < 8 else 8))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "mp_shard_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    log = os.environ.get("MRL_TEST_LOG")  # keep the worker's output (profiles/ wants the world-8 log)
    if log:
        with open(log, "w") as f:
            f.write(f"$ {' '.join(cmd)}\nexit code {r.returncode}\n--- stdout ---\n{r.stdout}\n--- stderr (tail) ---\n{r.stderr[-4000:]}\n")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert f"SHARD_CHECK OK world {world}" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    if world > 1:
        assert "SHARD_TIMEOUT ok" in r.stdout, r.stdout[-2000:]
