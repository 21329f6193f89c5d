"""Where do the microseconds of the sharded C2 step go?  (run under torchrun, or alone for world 1)

Times, with bench.py's flush-before-every-iteration discipline, the C2 step (sample 256 + update 256) as:
  plain        un-sharded calls
  p2p          mrl_prb_sample_sharded_p2p + update, statistics published by the update kernel (default protocol)
  p2p_sample   round-1 protocol
and the two kernels of each variant alone.  Prints one JSON line per rank 0.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    class A:
        pass
    ctx = bench.Ctx(A())
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    st = ctx.stream
    batch = int(os.environ.get("PROBE_BATCH", 256))
    steps = int(os.environ.get("PROBE_STEPS", 200))
    out = {"world": ctx.world, "batch": batch, "steps": steps}
    for mode in ("plain", "p2p", "p2p_sample"):
        h, cols, _ = bench.make_prb(ctx, 1_000_000, bench.C2_ROWS, 7)
        idx = torch.empty(batch, dtype=torch.int64, device="cuda")
        w = torch.empty(batch, dtype=torch.float32, device="cuda")
        outs = [torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in bench.C2_ROWS]
        tab = ffi.ptr_table([o.data_ptr() for o in outs])
        prio = torch.rand((64, batch), device="cuda") + 1e-3
        pp = [prio[i].data_ptr() for i in range(64)]
        if mode != "plain":
            buf = C.create_string_buffer(64)
            ffi.check(L.mrl_prb_shard_init(h, ctx.rank, ctx.world, buf))
            ffi.check(L.mrl_prb_shard_set_mode(h, 0 if mode == "p2p_sample" else 1))
            if ctx.world > 1:
                parts = [None] * ctx.world
                ctx.dist.all_gather_object(parts, buf.raw)
                allh = b"".join(parts)
            else:
                allh = buf.raw
            ffi.check(L.mrl_prb_shard_connect(h, C.create_string_buffer(allh, len(allh)), st))
            torch.cuda.synchronize()
            ctx.barrier()

        def sample(i):
            if mode == "plain":
                ffi.check(L.mrl_prb_sample(h, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, st))
            else:
                ffi.check(L.mrl_prb_sample_sharded_p2p(h, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, st))

        def update(i):
            ffi.check(L.mrl_prb_update(h, idx.data_ptr(), pp[i % 64], batch, st))

        def step(i):
            sample(i)
            update(i)

        res = {}
        for name, fn in (("step", step), ("sample_only", sample), ("update_only", update), ("step_again", step)):
            if name == "sample_only" and mode == "p2p":
                # consecutive samples re-publish through an extra tail kernel: time sample+update pairs instead
                continue
            ms, _ = ctx.timed(fn, steps, 10)
            res[name] = ms / steps * 1e3
        if mode != "plain":
            v = [C.c_int64() for _ in range(4)]
            ffi.check(L.mrl_prb_shard_diag(h, *[C.byref(x) for x in v], 0, st))
            res["diag"] = [x.value for x in v]
        out[mode] = res
        ffi.check(L.mrl_prb_destroy(h))
    if ctx.rank == 0:
        print(json.dumps(out))
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
