"""torchrun worker of tests/test_gpu_multi.py: sharded PER over N GPUs of one node.

Three exchanges of the root statistics -- peer mailboxes written by the mutating kernels ('p2p', default),
peer mailboxes written by the sample kernel ('p2p_sample', round 1) and the NCCL all-gather ('nccl') -- run the
same SPMD program and must agree with each other and with the CPU oracle bit for bit: a program with
consecutive samples, single-row and batched inserts between samples, several updates between samples and the
C5 batch (4096 per rank).  Then the bounded poll: a sample that no peer matches must give up within the
timeout and surface MRL_ERR_STATE instead of hanging the GPU."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mindrl_b200 as M  # noqa: E402
from mindrl_b200 import _ffi  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    _ffi.check(_ffi.lib().mrl_set_device(int(os.environ["LOCAL_RANK"])))
    shapes, types = [(4,), (1,)], [np.float32, np.int32]
    ok = True
    notes = []

    def expect(cond, what):
        nonlocal ok
        if not cond:
            ok = False
            notes.append(what)

    def dev(a):
        return torch.as_tensor(np.ascontiguousarray(a)).cuda()

    for cap, batch, steps in ((4096, 64, 10), (1 << 20, 4096, 3)):
        rng = np.random.default_rng(100 + rank)
        fill = cap - 100 * rank - 50  # ranks differ in size and mass; room left for the inserts below
        rows = [rng.normal(size=(fill, 4)).astype(np.float32), np.arange(fill, dtype=np.int32)[:, None]]
        pr = (rng.random(fill) * (rank + 1) + 0.01).astype(np.float32)

        def make():
            g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 3, rank)
            g.insert(*[dev(c) for c in rows])
            g.update_priorities(torch.arange(fill).cuda(), dev(pr))
            return g

        bufs = {k: make() for k in ("p2p", "p2p_sample", "nccl")}
        o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 3, rank)
        o.insert_batch(rows)
        o.update_priorities(np.arange(fill), pr)
        sh = {k: M.ShardedPriorityReplayBuffer(b, cap, exchange=k) for k, b in bufs.items()}
        expect(sh["p2p"].p2p and sh["p2p_sample"].p2p and not sh["nccl"].p2p, "exchange selection")
        auto = M.ShardedPriorityReplayBuffer(make(), cap, exchange="auto")
        expect(auto.exchange == "p2p", "auto did not pick the peer mailboxes on one node")

        def sample_all(beta, u):
            """sample on all three buffers and on the oracle (global statistics through NCCL)"""
            outs = {k: s.sample(beta, uniforms=None if u is None else dev(u)) for k, s in sh.items()}
            g = sh["nccl"].global_stats().cpu().numpy()
            want = o.sample(beta, uniforms=u, gstats=g)
            for k, a in outs.items():
                expect(np.array_equal(a[0].cpu().numpy(), want[0] + rank * cap), f"{k}: indices (cap {cap})")
                expect(np.array_equal(a[1].cpu().numpy().view(np.uint32), want[1].view(np.uint32)), f"{k}: weights (cap {cap})")
                expect(np.array_equal(a[2].cpu().numpy(), want[2]) and np.array_equal(a[3].cpu().numpy(), want[3]), f"{k}: rows")
            return outs, want

        def update_all(outs, want, newp):
            for k, s in sh.items():
                s.update_priorities(outs[k][0], dev(newp))
            o.update_priorities(want[0], newp)

        for step in range(steps):
            u = rng.random(batch).astype(np.float32)
            if step % 2:  # the optional early wait for the peers (also in front of consecutive samples below)
                for s in sh.values():
                    s.wait_for_peers()
            outs, want = sample_all(0.4, u)
            if step % 3 == 1:  # a second sample with nothing in between: the statistics are re-published
                if step % 2:
                    sh["p2p"].wait_for_peers()
                outs, want = sample_all(0.5, rng.random(batch).astype(np.float32))
            update_all(outs, want, (rng.random(batch) * 2 + 0.01).astype(np.float32))
            if step % 2 == 0:  # actors keep inserting between the learner's samples: single rows and a batch
                one = [rng.normal(size=(4,)).astype(np.float32), np.array([step], np.int32)]
                few = [rng.normal(size=(5, 4)).astype(np.float32), np.arange(5, dtype=np.int32)[:, None]]
                for s in sh.values():
                    s.insert(*[dev(c) for c in one])
                    s.insert(*[dev(c) for c in few])
                o.insert(*one)
                o.insert_batch(few)
            if step % 4 == 3:  # two updates between two samples
                idx = rng.integers(0, fill, size=batch)
                newp = (rng.random(batch) + 0.05).astype(np.float32)
                for s in sh.values():
                    s.local.update_priorities(dev(idx), dev(newp))
                o.update_priorities(idx, newp)
        # weights against the true per-shard sampling probability (option "shard_weights"), all three exchanges
        _ffi.set_option("shard_weights", 1)
        O.set_shard_weights(1)
        outs, want = sample_all(0.6, rng.random(batch).astype(np.float32))
        _ffi.set_option("shard_weights", 0)
        O.set_shard_weights(0)
        update_all(outs, want, (rng.random(batch) + 0.05).astype(np.float32))
        for k, b in bufs.items():
            gs, gm = b.tree()
            os_, om = o.tree()
            expect(np.array_equal(gs[1:], os_[1:]) and np.array_equal(gm[1:], om[1:]), f"{k}: tree (cap {cap})")
        d = sh["p2p"].exchange_diag()
        expect(d["timeouts"] == 0 and d["polled_calls"] > 0, f"p2p diag {d}")
        d0 = sh["p2p_sample"].exchange_diag()
        expect(d0["timeouts"] == 0, f"p2p_sample diag {d0}")
        if rank == 0:
            print(f"SHARD_DIAG cap {cap} batch {batch}: publish_on_mutation {d} | publish_on_sample {d0}", flush=True)
        torch.cuda.synchronize()
        dist.barrier()

    # ---- the poll is bounded ------------------------------------------------------------------------
    _ffi.set_option("shard_timeout_ms", 300)
    g = M.PriorityReplayBuffer(0.6, 4096, 64, shapes, types, 3, rank, carrier="numpy")
    g.insert(np.zeros((1000, 4), np.float32), np.zeros((1000, 1), np.int32))
    shp = M.ShardedPriorityReplayBuffer(g, 4096, exchange="p2p")
    L = _ffi.lib()
    import ctypes as C
    idx, w = np.empty(64, np.int64), np.empty(64, np.float32)
    rc = L.mrl_prb_sample_sharded_p2p_host(g._handle, 64, 0.4, None, idx.ctypes.data, w.ctypes.data, None, 0)
    expect(rc == 0 and np.all(np.isfinite(w)), "matched sample before the timeout test")
    torch.cuda.synchronize()
    dist.barrier()
    if world > 1 and rank == 0:  # a call no peer matches
        t0 = time.time()
        rc = L.mrl_prb_sample_sharded_p2p_host(g._handle, 64, 0.4, None, idx.ctypes.data, w.ctypes.data, None, 0)
        dt = time.time() - t0
        expect(rc == _ffi.MRL_ERR_STATE, f"unmatched sample returned {rc}, not MRL_ERR_STATE")
        expect(0.25 < dt < 5.0, f"unmatched sample took {dt:.2f}s (timeout 0.3s)")
        expect(b"rank" in L.mrl_last_error(), "error text names the missing rank")
        rc2 = L.mrl_prb_sample_sharded_p2p_host(g._handle, 64, 0.4, None, idx.ctypes.data, w.ctypes.data, None, 0)
        expect(rc2 == _ffi.MRL_ERR_STATE, "the handle stays in the error state")
        print(f"SHARD_TIMEOUT ok: gave up after {dt:.2f}s", flush=True)
    _ffi.set_option("shard_timeout_ms", 10000)
    torch.cuda.synchronize()
    dist.barrier()

    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not ok:
        print(f"rank {rank} FAILED: {notes[:10]}", flush=True)
    if rank == 0:
        print("SHARD_CHECK", "OK" if int(flag.item()) == 1 else "FAIL", "world", world, flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
