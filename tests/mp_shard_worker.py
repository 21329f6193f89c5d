"""torchrun worker of tests/test_gpu_multi.py: sharded PER over N GPUs, peer-mailbox exchange against
the NCCL all-gather path and against the CPU oracle (same seeded inputs on both sides)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mindrl_b200 as M  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    from mindrl_b200 import _ffi
    _ffi.check(_ffi.lib().mrl_set_device(int(os.environ["LOCAL_RANK"])))
    cap, batch = 4096, 64
    shapes, types = [(4,), (1,)], [np.float32, np.int32]
    rng = np.random.default_rng(100 + rank)
    rows = [rng.normal(size=(cap - 100 * rank, 4)).astype(np.float32),
            np.arange(cap - 100 * rank, dtype=np.int32)[:, None]]
    pr = (rng.random(cap - 100 * rank) * (rank + 1) + 0.01).astype(np.float32)

    def make():
        g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 3, rank)
        g.insert(*[torch.as_tensor(c).cuda() for c in rows])
        g.update_priorities(torch.arange(len(pr)).cuda(), torch.as_tensor(pr).cuda())
        return g

    g_p2p, g_nccl = make(), make()
    o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 3, rank)
    o.insert_batch(rows)
    o.update_priorities(np.arange(len(pr)), pr)
    sh_p2p = M.ShardedPriorityReplayBuffer(g_p2p, cap, exchange="p2p")
    sh_nccl = M.ShardedPriorityReplayBuffer(g_nccl, cap, exchange="nccl")
    assert sh_p2p.p2p and not sh_nccl.p2p
    ok = True
    for step in range(12):
        u = rng.random(batch).astype(np.float32)
        a = sh_p2p.sample(0.4, uniforms=torch.as_tensor(u).cuda())
        b = sh_nccl.sample(0.4, uniforms=torch.as_tensor(u).cuda())
        gstats = sh_nccl.global_stats().cpu().numpy()
        want = o.sample(0.4, uniforms=u, gstats=gstats)
        for x, y in zip(a, b):
            ok &= bool(torch.equal(x, y))
        ok &= np.array_equal(a[0].cpu().numpy(), want[0] + rank * cap)
        ok &= np.array_equal(a[1].cpu().numpy().view(np.uint32), want[1].view(np.uint32))
        ok &= np.array_equal(a[3].cpu().numpy(), want[3])
        newp = (rng.random(batch) * 2 + 0.01).astype(np.float32)
        sh_p2p.update_priorities(a[0], torch.as_tensor(newp).cuda())
        sh_nccl.update_priorities(b[0], torch.as_tensor(newp).cuda())
        o.update_priorities(want[0], newp)
    gs, gm = g_p2p.tree()
    os_, om = o.tree()
    ok &= np.array_equal(gs[1:], os_[1:]) and np.array_equal(gm[1:], om[1:])
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARD_CHECK", "OK" if int(flag.item()) == 1 else "FAIL", "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
