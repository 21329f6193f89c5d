"""parity + timing of the TMA-ring scan kernels (scan_tma.cu) against the oracle and the register kernels"""
import os, sys, json
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindrl_b200 as M
from mindrl_b200 import _ffi
from oracle import oracle as O

dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
fails = 0


def close(got, want, what):
    global fails
    got = got.cpu().numpy()
    tol = 1e-5 * max(1.0, float(np.abs(want).max()))
    err = float(np.abs(got - want).max()) if want.size else 0.0
    if not (err <= tol) or not np.isfinite(got).all():
        fails += 1
        bad = np.argwhere(~(np.abs(got - want) <= tol))
        print("FAIL", what, "err", err, "tol", tol, "first bad", bad[:3].tolist(), flush=True)


def parity_tm(cols, rpu=8, groups=1):
    _ffi.set_option("tm_kernel", 6)
    _ffi.set_option("strip_groups", groups)
    _ffi.set_option("strip_cols", cols)
    _ffi.set_option("strip_rpu", rpu)
    for T, B in [(1, 16), (5, 16), (64, 32), (65, 48), (127, 80), (128, 64), (300, 4112), (1000, 96), (2048, 128),
                 (513, 36), (77, 4)]:
        rng = np.random.default_rng(T * 7 + B)
        r = rng.normal(size=(T, B)).astype(np.float32)
        val = rng.normal(size=(T + 1, B)).astype(np.float32)
        lv = rng.normal(size=(B,)).astype(np.float32)
        for with_done in (False, True):
            if with_done and B % 16:
                continue
            d = (rng.random((T, B)) < 0.05) if with_done else None
            dd = None if d is None else dev(d)
            adv_r, ret_r = O.gae(r, val[:T], None, val[T], d, 0.99, 0.95, layout=0)
            vd = dev(val)
            adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=vd[T], done=dd, mode="parallel")
            close(adv, adv_r, f"tm{cols} gae adv T{T} B{B} d{with_done}")
            close(ret, ret_r, f"tm{cols} gae ret T{T} B{B} d{with_done}")
            adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=None, done=dd, mode="parallel", want_returns=False)
            adv_r, _ = O.gae(r, val[:T], None, None, d, 0.99, 0.95, layout=0)
            close(adv, adv_r, f"tm{cols} gae nolast T{T} B{B} d{with_done}")
            ref = O.discounted_return(r, d, lv, 0.99)
            if d is not None:
                close(M.DiscountedReturn(0.99, mode="parallel")(dev(r), dd, dev(lv)), ref, f"tm{cols} DR T{T} B{B}")
            else:
                ref = O.discounted_return(r, None, lv, 0.9)
                close(M.discounted_reward(dev(r), 0.9, last_value=dev(lv), layout="time_major", mode="parallel"), ref,
                      f"tm{cols} DR nodone T{T} B{B}")
        b = rng.uniform(0.0, 1.0, size=(T, B)).astype(np.float32)
        ref = O.affine_scan(r, b, lv, layout=0)
        close(M.affine_scan(dev(r), dev(b), dev(lv), layout="time_major", mode="parallel"), ref, f"tm{cols} affine T{T} B{B}")
    _ffi.set_option("tm_kernel", 0)
    _ffi.set_option("strip_cols", 0)
    _ffi.set_option("strip_rpu", 0)
    _ffi.set_option("strip_groups", 0)


def parity_bm():
    _ffi.set_option("bm_kernel", 2)
    for B, T in [(1, 4), (3, 16), (7, 128), (30, 1000), (17, 144), (64, 2048), (5, 4096), (9, 2064), (2, 4112),
                 (300, 512), (3, 6160)]:
        rng = np.random.default_rng(B * 1009 + T)
        r = rng.normal(size=(B, T)).astype(np.float32)
        v = rng.normal(size=(B, T)).astype(np.float32)
        nv = rng.normal(size=(B, T)).astype(np.float32)
        lv = rng.normal(size=(B,)).astype(np.float32)
        for with_done in (False, True):
            if with_done and T % 16:
                continue
            d = (rng.random((B, T)) < 0.05) if with_done else None
            dd = None if d is None else dev(d)
            ref = O.discounted_return(r, d, lv, 0.99, layout=1)
            close(M.discounted_reward(dev(r), 0.99, last_value=dev(lv), done=dd), ref, f"bm DR B{B} T{T} d{with_done}")
            adv_r, ret_r = O.gae(r, v, None, lv, d, 0.99, 0.95, layout=1)
            adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, last_value=dev(lv), done=dd, layout="batch_major")
            close(adv, adv_r, f"bm gae adv B{B} T{T} d{with_done}")
            close(ret, ret_r, f"bm gae ret B{B} T{T} d{with_done}")
            adv_r, ret_r = O.gae(r, v, nv, None, d, 0.99, 0.95, layout=1)
            adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, next_value=dev(nv), done=dd, layout="batch_major")
            close(adv, adv_r, f"bm gae nv adv B{B} T{T} d{with_done}")
            close(ret, ret_r, f"bm gae nv ret B{B} T{T} d{with_done}")
        b = rng.uniform(0.0, 1.0, size=(B, T)).astype(np.float32)
        ref = O.affine_scan(r, b, lv, layout=1)
        close(M.affine_scan(dev(r), dev(b), dev(lv), layout="batch_major"), ref, f"bm affine B{B} T{T}")
    _ffi.set_option("bm_kernel", 0)


def timeit(fns, reps=10):
    """fns: the same call on different buffer sets (rotated so that nothing is reused out of L2);
    back-to-back launches inside one event pair -> average device time per launch"""
    for f in fns:
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for f in fns:
                f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (reps * len(fns)))
    return round(best, 2)


def timing():
    L = _ffi.lib()
    T, B = 2048, 4096
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    p = lambda x: x.data_ptr()
    sets = []
    for i in range(5):
        r = torch.randn((T, B), device="cuda", generator=g)
        v = torch.randn((T + 1, B), device="cuda", generator=g)
        d = (torch.rand((T, B), device="cuda", generator=g) < 1 / 200).to(torch.uint8)
        adv = torch.empty((T, B), device="cuda"); ret = torch.empty((T, B), device="cuda")
        sets.append((r, v, d, adv, ret))
    mk = lambda f: [(lambda s=s: f(*s)) for s in sets]
    gae_tm = mk(lambda r, v, d, adv, ret: L.mrl_gae(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), T, B, 0.99, 0.95, 0, 0, 0))
    dr_tm = mk(lambda r, v, d, adv, ret: L.mrl_discounted_return(p(r), p(d), p(v[T]), p(adv), T, B, 1, 0.99, 0, 0, 0))
    gae_bm = mk(lambda r, v, d, adv, ret: L.mrl_gae(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), 2048, 4096, 0.99, 0.95, 1, 0, 0))
    dr_bm = mk(lambda r, v, d, adv, ret: L.mrl_discounted_return(p(r), p(d), p(v[T]), p(adv), 2048, 4096, 1, 0.99, 1, 0, 0))
    res = {}
    base = {"tm_kernel": 6, "strip_cols": 32, "strip_rpu": 16, "strip_groups": 0, "strip_stages": 0, "strip_min_cols": 2048}
    for name, opts in [("tm_strip32x16_g2", {"strip_groups": 2}), ("tm_strip32x16_g2_s3", {"strip_groups": 2, "strip_stages": 3}),
                       ("tm_strip32x16_g2_s4", {"strip_groups": 2, "strip_stages": 4}),
                       ("tm_strip32x16_g1", {"strip_groups": 1}), ("tm_strip32x16_g1_s3", {"strip_groups": 1, "strip_stages": 3}),
                       ("tm_strip32x8", {"strip_rpu": 8}), ("tm_strip16x8", {"strip_cols": 16, "strip_rpu": 8}),
                       ("tm_strip16x16", {"strip_cols": 16}),
                       ("tm_pipe", {"tm_kernel": 0, "strip_min_cols": 1 << 40})]:
        o = dict(base); o.update(opts)
        for k_, v_ in o.items():
            _ffi.set_option(k_, v_)
        res[name] = {"gae": timeit(gae_tm), "dr": timeit(dr_tm)}
        print(name, res[name], flush=True)
    for k_, v_ in base.items():
        _ffi.set_option(k_, 0 if k_ != "strip_min_cols" else 2048)
    for name, mode, per_sm in [("bm_row_1", 2, 1), ("bm_row_2", 2, 2), ("bm_row_3", 2, 3), ("bm_row_4", 2, 4), ("bm_row_5", 2, 5), ("bm_warp", 1, 3)]:
        _ffi.set_option("bm_kernel", mode)
        _ffi.set_option("row_ctas_per_sm", per_sm)
        res[name] = {"gae": timeit(gae_bm), "dr": timeit(dr_bm)}
        print(name, res[name], flush=True)
    _ffi.set_option("bm_kernel", 0); _ffi.set_option("row_ctas_per_sm", 3)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/tma_scan_timing.json", "w"), indent=1)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("all", "parity"):
        parity_tm(32, 16, 2); print("tm 32 16 g2 done", fails, flush=True)
        for cols in (32, 16):
            for rpu in (8, 16):
                parity_tm(cols, rpu); print("tm", cols, rpu, "done", fails, flush=True)
        parity_bm(); print("bm done", fails, flush=True)
    if what in ("all", "timing"):
        timing()
    print("FAILS", fails)
    sys.exit(1 if fails else 0)
