"""Where do the microseconds of the end-to-end C2 step go?  Host clock around each *_host call, with and without
the L2 flush + idle GPU that bench.py's discipline puts in front of every step."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench


def main():
    class A: pass
    ctx = bench.Ctx(A())
    torch, L, ffi = ctx.torch, ctx.L, ctx.ffi
    st = ctx.stream
    batch = 256
    h, cols, _ = bench.make_prb(ctx, 1_000_000, bench.C2_ROWS, 7)
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_idx, h_w = pin((batch,), torch.int64), pin((batch,), torch.float32)
    h_out = [pin((batch, r), torch.uint8) for r in bench.C2_ROWS]
    tab = ffi.ptr_table([o.data_ptr() for o in h_out])
    prio = pin((64, batch), torch.float32); prio.copy_(torch.rand(64, batch) + 1e-3)
    res = {}
    for flush in (True, False):
        ts, tu, ty = [], [], []
        for i in range(300):
            if flush:
                ctx.flush(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            ffi.check(L.mrl_prb_sample_host(h, batch, 0.4, None, h_idx.data_ptr(), h_w.data_ptr(), tab, st))
            t1 = time.perf_counter()
            ffi.check(L.mrl_prb_update_host(h, h_idx.data_ptr(), prio[i % 64].data_ptr(), batch, st))
            t2 = time.perf_counter()
            ffi.check(L.mrl_stream_sync(st))
            t3 = time.perf_counter()
            if i >= 50:
                ts.append(t1 - t0); tu.append(t2 - t1); ty.append(t3 - t2)
        med = lambda x: float(np.median(x)) * 1e6
        res["flush" if flush else "back_to_back"] = {"sample_host_us": med(ts), "update_host_call_us": med(tu),
                                                      "drain_us": med(ty), "step_us": med(ts) + med(tu) + med(ty)}
    # an empty host call for scale: ctypes overhead
    t0 = time.perf_counter()
    for _ in range(10000):
        L.mrl_launch_count()
    res["ctypes_call_us"] = (time.perf_counter() - t0) / 10000 * 1e6
    print(res)

main()
