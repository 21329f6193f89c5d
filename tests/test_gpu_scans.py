"""GPU parity of the reverse scans against the CPU oracle, the reference's known answers and the
golden vectors.

Tolerances.  Sequential mode: bit-exact.  Parallel modes re-associate fp32, so they are held to
1e-5 RELATIVE TO THE MAGNITUDE OF THE ARRAY: |gpu - oracle| <= 1e-5 * max(||oracle||_inf, 1).  An
element-wise relative bound is not meaningful here: a return is a sum of ~1/(1-gamma) terms, its
fp32 rounding error scales with the partial sums, and elements that cancel to ~0 carry the same
absolute error as their neighbours.  test_full_size_c4_against_oracle_and_float64 shows that the fp32
oracle itself is no closer than that to a float64 evaluation, and that the parallel kernels are as
close to float64 as the oracle is."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import oracle as O  # noqa: E402

import mindrl_b200 as M  # noqa: E402

RTOL = 1e-5


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def close(gpu, ref):
    gpu = gpu.cpu().numpy() if hasattr(gpu, "cpu") else np.asarray(gpu)
    err = np.abs(gpu.astype(np.float64) - ref.astype(np.float64))
    assert gpu.shape == ref.shape
    if ref.size == 0:
        return
    bound = RTOL * max(float(np.abs(ref).max()), 1.0)
    assert err.max() <= bound, f"max err {err.max()} (bound {bound}) at {np.unravel_index(err.argmax(), err.shape)}"


def exact(gpu, ref):
    gpu = gpu.cpu().numpy() if hasattr(gpu, "cpu") else np.asarray(gpu)
    assert gpu.shape == ref.shape
    assert np.array_equal(gpu.view(np.uint32), ref.view(np.uint32))


def test_known_answers_device_and_host():
    net = M.DiscountedReturn(gamma=0.99)
    r = np.array([[1], [1], [1], [1]], np.float32)
    d = np.array([[False], [False], [True], [False]])
    want = np.array([[2.9701], [1.99], [1.0], [2.98]], np.float32)
    out = net(dev(r), dev(d), dev(np.array([2.0], np.float32)))
    assert out.is_cuda and np.allclose(out.cpu().numpy(), want)
    out = net(r, d, np.array([2.0], np.float32))  # numpy carrier -> host entry point
    assert isinstance(out, np.ndarray) and np.allclose(out, want)
    r = np.array([1, 1, 1, 1], np.float32)
    d = np.array([False, True, False, True])
    want = np.array([1.99, 1, 1.99, 1.0], np.float32)
    assert np.allclose(net(dev(r), dev(d), dev(np.float32(2.0))).cpu().numpy(), want)
    assert np.allclose(net(r, d, np.float32(2.0)), want)
    with pytest.raises(ValueError):
        M.DiscountedReturn(gamma=1.5)


def test_discounted_return_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "discounted_return_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        args = (dev(g[n + "/reward"]), dev(g[n + "/done"]), dev(g[n + "/last"]))
        gamma = float(g[n + "/gamma"])
        exact(M.DiscountedReturn(gamma, mode="sequential")(*args), g[n + "/out"])
        close(M.DiscountedReturn(gamma, mode="parallel")(*args), g[n + "/out"])
        close(M.DiscountedReturn(gamma)(*args), g[n + "/out"])


def test_ppo_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "ppo_scans_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        gamma, lam = float(g[n + "/gamma"]), float(g[n + "/lam"])
        r, v, nv = dev(g[n + "/reward"]), dev(g[n + "/value"]), dev(g[n + "/next_value"])
        exact(M.discounted_reward(r, gamma, layout="batch_major", mode="sequential"), g[n + "/discounted_r"])
        close(M.discounted_reward(r, gamma, layout="batch_major"), g[n + "/discounted_r"])
        adv, ret = M.gae(r, v, gamma, lam, next_value=nv, layout="batch_major", mode="sequential")
        exact(adv, g[n + "/advantage"])
        adv, ret = M.gae(r, v, gamma, lam, next_value=nv, layout="batch_major")
        close(adv, g[n + "/advantage"])
        close(ret, g[n + "/advantage"] + g[n + "/value"])


@pytest.mark.parametrize("T,B,E", [(1, 1, 1), (5, 3, 1), (64, 33, 1), (257, 40, 3), (1000, 30, 1),
                                   (2048, 96, 1), (300, 5000, 1), (513, 7, 2)])
def test_discounted_return_time_major(T, B, E):
    rng = np.random.default_rng(T * 131 + B)
    shape = (T, B, E) if E > 1 else (T, B)
    r = rng.normal(size=shape).astype(np.float32)
    d = rng.random((T, B)) < 0.05
    lv = rng.normal(size=shape[1:]).astype(np.float32)
    ref = O.discounted_return(r, d, lv, 0.99)
    exact(M.DiscountedReturn(0.99, mode="sequential")(dev(r), dev(d), dev(lv)), ref)
    close(M.DiscountedReturn(0.99, mode="parallel")(dev(r), dev(d), dev(lv)), ref)
    close(M.DiscountedReturn(0.99)(dev(r), dev(d), dev(lv)), ref)
    # no done flags, no bootstrap
    ref = O.discounted_return(r, None, None, 0.9)
    close(M.discounted_reward(dev(r.reshape(T, -1)), 0.9, layout="time_major"), ref.reshape(T, -1))


@pytest.mark.parametrize("B,T", [(1, 1), (3, 5), (7, 128), (30, 1000), (17, 130), (64, 2048), (5, 127),
                                 (9, 4), (2, 4099)])
@pytest.mark.parametrize("with_done", [False, True])
def test_batch_major(B, T, with_done):
    rng = np.random.default_rng(B * 1009 + T)
    r = rng.normal(size=(B, T)).astype(np.float32)
    v = rng.normal(size=(B, T)).astype(np.float32)
    nv = rng.normal(size=(B, T)).astype(np.float32)
    lv = rng.normal(size=(B,)).astype(np.float32)
    d = (rng.random((B, T)) < 0.05) if with_done else None
    dd = None if d is None else dev(d)
    ref = O.discounted_return(r, d, lv, 0.99, layout=1)
    close(M.discounted_reward(dev(r), 0.99, last_value=dev(lv), done=dd), ref)
    exact(M.discounted_reward(dev(r), 0.99, last_value=dev(lv), done=dd, mode="sequential"), ref)
    # GAE, bootstrap from value[t+1] / last_value
    adv_r, ret_r = O.gae(r, v, None, lv, d, 0.99, 0.95, layout=1)
    adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, last_value=dev(lv), done=dd, layout="batch_major")
    close(adv, adv_r)
    close(ret, ret_r)
    adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, last_value=dev(lv), done=dd, layout="batch_major",
                     mode="sequential")
    exact(adv, adv_r)
    exact(ret, ret_r)
    # GAE, PPO form with a separate next_value array
    adv_r, ret_r = O.gae(r, v, nv, None, d, 0.99, 0.95, layout=1)
    adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, next_value=dev(nv), done=dd, layout="batch_major")
    close(adv, adv_r)
    close(ret, ret_r)


@pytest.mark.parametrize("T,B", [(1, 1), (25, 8), (64, 33), (1000, 30), (2048, 128), (100, 6000)])
@pytest.mark.parametrize("with_done", [False, True])
def test_gae_time_major(T, B, with_done):
    rng = np.random.default_rng(T * 7 + B)
    r = rng.normal(size=(T, B)).astype(np.float32)
    val = rng.normal(size=(T + 1, B)).astype(np.float32)
    d = (rng.random((T, B)) < 0.05) if with_done else None
    dd = None if d is None else dev(d)
    adv_r, ret_r = O.gae(r, val[:T], None, val[T], d, 0.99, 0.95, layout=0)
    vd = dev(val)
    for mode, cmp in (("sequential", exact), ("parallel", close), ("auto", close)):
        adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=vd[T], done=dd, mode=mode)
        cmp(adv, adv_r)
        cmp(ret, ret_r)
    adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=vd[T], done=dd, want_returns=False)
    assert ret is None
    close(adv, adv_r)


def test_gae_host_path():
    rng = np.random.default_rng(11)
    T, B = 300, 64
    r = rng.normal(size=(T, B)).astype(np.float32)
    v = rng.normal(size=(T, B)).astype(np.float32)
    lv = rng.normal(size=(B,)).astype(np.float32)
    d = rng.random((T, B)) < 0.05
    adv_r, ret_r = O.gae(r, v, None, lv, d, 0.99, 0.95, layout=0)
    adv, ret = M.gae(r, v, 0.99, 0.95, last_value=lv, done=d)
    assert isinstance(adv, np.ndarray)
    close(adv, adv_r)
    close(ret, ret_r)


def test_affine_scan_vtrace_form():
    rng = np.random.default_rng(12)
    for layout, shape in (("time_major", (80, 40)), ("batch_major", (40, 80)), ("batch_major", (3, 77))):
        a = rng.normal(size=shape).astype(np.float32)
        b = (0.99 * np.minimum(1.0, rng.random(shape) * 2) * (rng.random(shape) > 0.1)).astype(np.float32)
        lay = 0 if layout == "time_major" else 1
        xl = rng.normal(size=(shape[1] if lay == 0 else shape[0],)).astype(np.float32)
        ref = O.affine_scan(a, b, xl, layout=lay)
        close(M.affine_scan(dev(a), dev(b), dev(xl), layout=layout), ref)
        exact(M.affine_scan(dev(a), dev(b), dev(xl), layout=layout, mode="sequential"), ref)


def test_vtrace_golden(golden_dir):
    """get_vs_and_advantages against vectors produced by the reference's own vtrace.py"""
    g = np.load(os.path.join(golden_dir, "vtrace_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        args = [dev(g[f"{n}/{k}"]) for k in ("log_rhos", "discounts", "rewards", "values", "boot")]
        for mode in ("auto", "sequential"):
            vs, pg = M.get_vs_and_advantages(*args, 1.0, 1.0, 1.0, dev(g[f"{n}/masks"]), mode=mode)
            close(vs, g[f"{n}/vs"])   # exp() of the carrier differs from numpy's in the last ulp
            close(pg, g[f"{n}/pg"])


def test_full_size_c4_against_oracle_and_float64():
    """BASELINE config 4: 4096 envs x 2048 steps, gamma .99 lambda .95, random done flags"""
    rng = np.random.default_rng(4)
    T, B = 2048, 4096
    r = rng.normal(size=(T, B)).astype(np.float32)
    val = rng.normal(size=(T + 1, B)).astype(np.float32)
    d = rng.random((T, B)) < 1 / 200
    adv_r, ret_r = O.gae(r, val[:T], None, val[T], d, 0.99, 0.95, layout=0)
    vd = dev(val)
    adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=vd[T], done=dev(d))
    close(adv, adv_r)
    close(ret, ret_r)
    adv_s, _ = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=vd[T], done=dev(d), mode="sequential")
    exact(adv_s, adv_r)
    # float64 ground truth: the parallel kernel is no further from it than the fp32 oracle (x4)
    m = 1.0 - d.astype(np.float64)
    delta = r.astype(np.float64) + np.float32(0.99).astype(np.float64) * np.concatenate([val[1:T], val[T:]]).astype(np.float64) * m - val[:T].astype(np.float64)
    gl = np.float64(np.float32(0.99) * np.float32(0.95))
    truth = np.zeros((T, B))
    acc = np.zeros(B)
    for t in range(T - 1, -1, -1):
        acc = delta[t] + gl * m[t] * acc
        truth[t] = acc
    e_gpu = np.abs(adv.cpu().numpy() - truth).max()
    e_orc = np.abs(adv_r - truth).max()
    assert e_gpu <= 2 * e_orc + 1e-6
    # the fp32 oracle itself misses an element-wise 1e-5 relative bound against float64
    rel_orc = (np.abs(adv_r - truth) / np.maximum(np.abs(truth), 1e-30)).max()
    assert rel_orc > 1e-5
    # batch-major twin of the same problem
    adv_b, ret_b = M.gae(dev(r.T.copy()), dev(val[:T].T.copy()), 0.99, 0.95, last_value=vd[T],
                         done=dev(d.T.copy()), layout="batch_major")
    close(adv_b, adv_r.T)
    close(ret_b, ret_r.T)
    assert np.abs(adv_b.cpu().numpy().T - truth).max() <= 2 * e_orc + 1e-6
    dr_r = O.discounted_return(r, d, val[T], 0.99)
    close(M.DiscountedReturn(0.99)(dev(r), dev(d), vd[T]), dr_r)
    exact(M.DiscountedReturn(0.99, mode="sequential")(dev(r), dev(d), vd[T]), dr_r)


def test_empty_and_errors():
    out = M.DiscountedReturn(0.9)(torch.zeros((0, 4), device="cuda"), torch.zeros((0, 4), dtype=torch.bool, device="cuda"),
                                  torch.zeros((4,), device="cuda"))
    assert tuple(out.shape) == (0, 4)
    with pytest.raises(ValueError):
        M.DiscountedReturn(0.9)(torch.zeros((3, 4), device="cuda"), torch.zeros((3, 5), dtype=torch.bool, device="cuda"),
                                torch.zeros((4,), device="cuda"))


@pytest.mark.parametrize("n", [1, 7, 1000, 123457, 4096 * 2048])
def test_normalize_advantage(n):
    """ppo.py:361-368: (adv - mean) / (sqrt(var) + 1e-8) against a float64 evaluation"""
    rng = np.random.default_rng(n)
    x = (rng.normal(size=n) * 3 + 0.7).astype(np.float32)
    got = M.normalize_advantage(dev(x)).cpu().numpy()
    mean = x.astype(np.float64).mean()
    var = x.astype(np.float64).var()
    want = ((x - np.float32(mean)) * (np.float32(1.0) / (np.sqrt(np.float32(var)) + np.float32(1e-8)))).astype(np.float32)
    if n == 1:
        assert got[0] == 0.0
    else:
        assert np.allclose(got, want, rtol=2e-6, atol=2e-6)
        assert abs(got.astype(np.float64).mean()) < 1e-4 and abs(got.astype(np.float64).std() - 1) < 1e-3
    again = M.normalize_advantage(dev(x)).cpu().numpy()
    assert np.array_equal(got, again)  # fixed accumulation order


# ---- TMA-ring kernels (scan_tma.cu), forced through the library options ------------------------
class _options:
    """sets library options for the duration of a test and restores the defaults afterwards"""
    DEFAULTS = {"host_scan_chunks": 0, "tm_kernel": 0, "strip_cols": 0, "strip_rpu": 0, "strip_groups": 0, "strip_stages": 0,
                "bm_kernel": 0, "row_stages": 0, "row_spl": 8}

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        from mindrl_b200 import _ffi
        for k, v in self.kw.items():
            _ffi.set_option(k, v)

    def __exit__(self, *exc):
        from mindrl_b200 import _ffi
        for k in self.kw:
            _ffi.set_option(k, self.DEFAULTS[k])


TM_SHAPES = [(1, 16), (5, 16), (64, 32), (65, 48), (127, 80), (128, 64), (300, 4112), (1000, 96), (2048, 128),
             (513, 36), (77, 4), (40, 3552), (33, 3696)]


@pytest.mark.parametrize("cols,rpu,groups,stages", [(32, 16, 1, 0), (32, 16, 2, 0), (32, 16, 1, 2), (32, 8, 1, 0),
                                                    (16, 8, 1, 0), (16, 16, 1, 2), (32, 32, 1, 0), (32, 32, 1, 2), (28, 16, 1, 0), (28, 32, 1, 0), (28, 16, 1, 2)])
def test_tma_strip_kernels(cols, rpu, groups, stages):
    """time-major strips: ragged T (partial first tile), ragged N (clipped last strip), bootstrap row,
    with and without done flags, every recurrence; ring depth 2 forces stage reuse on the small shapes"""
    with _options(tm_kernel=6, strip_cols=cols, strip_rpu=rpu, strip_groups=groups, strip_stages=stages):
        for T, B in TM_SHAPES:
            rng = np.random.default_rng(T * 7 + B)
            r = rng.normal(size=(T, B)).astype(np.float32)
            val = rng.normal(size=(T + 1, B)).astype(np.float32)
            lv = rng.normal(size=(B,)).astype(np.float32)
            vd = dev(val)
            for with_done in (False, True):
                if with_done and B % 16:
                    continue  # (falls back to the register kernels: covered by test_gae_time_major)
                d = (rng.random((T, B)) < 0.05) if with_done else None
                dd = None if d is None else dev(d)
                adv_r, ret_r = O.gae(r, val[:T], None, val[T], d, 0.99, 0.95, layout=0)
                adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, last_value=vd[T], done=dd, mode="parallel")
                close(adv, adv_r)
                close(ret, ret_r)
                adv_r, _ = O.gae(r, val[:T], None, None, d, 0.99, 0.95, layout=0)
                adv, ret = M.gae(dev(r), vd[:T], 0.99, 0.95, done=dd, mode="parallel", want_returns=False)
                assert ret is None
                close(adv, adv_r)
                if with_done:
                    close(M.DiscountedReturn(0.99, mode="parallel")(dev(r), dd, dev(lv)), O.discounted_return(r, d, lv, 0.99))
                else:
                    close(M.discounted_reward(dev(r), 0.99, last_value=dev(lv), layout="time_major", mode="parallel"),
                          O.discounted_return(r, None, lv, 0.99))
                    close(M.discounted_reward(dev(r), 0.9, layout="time_major", mode="parallel"),
                          O.discounted_return(r, None, None, 0.9))
            b = rng.uniform(0.0, 1.0, size=(T, B)).astype(np.float32)
            close(M.affine_scan(dev(r), dev(b), dev(lv), layout="time_major", mode="parallel"),
                  O.affine_scan(r, b, lv, layout=0))


@pytest.mark.parametrize("spl", [8, 16])
@pytest.mark.parametrize("stages", [0, 2, 3])
def test_tma_row_kernel(stages, spl):
    """batch-major rows: several chunks per environment, a partial last chunk, T below one tile; 8 steps per lane
    (8 consumer warps) and 16 steps per lane (4 consumer warps)"""
    with _options(bm_kernel=2, row_stages=stages, row_spl=spl):
        for B, T in [(1, 4), (3, 16), (7, 128), (30, 1000), (17, 144), (64, 2048), (5, 4096), (9, 2064), (2, 4112),
                     (300, 512), (3, 6160), (8, 16384)]:
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
                close(M.discounted_reward(dev(r), 0.99, last_value=dev(lv), done=dd),
                      O.discounted_return(r, d, lv, 0.99, layout=1))
                adv_r, ret_r = O.gae(r, v, None, lv, d, 0.99, 0.95, layout=1)
                adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, last_value=dev(lv), done=dd, layout="batch_major")
                close(adv, adv_r)
                close(ret, ret_r)
                adv_r, ret_r = O.gae(r, v, nv, None, d, 0.99, 0.95, layout=1)
                adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, next_value=dev(nv), done=dd, layout="batch_major")
                close(adv, adv_r)
                close(ret, ret_r)
            b = rng.uniform(0.0, 1.0, size=(B, T)).astype(np.float32)
            close(M.affine_scan(dev(r), dev(b), dev(lv), layout="batch_major"), O.affine_scan(r, b, lv, layout=1))


@pytest.mark.parametrize("layout", ["time_major", "batch_major"])
def test_tma_ring_reuse_identity(layout):
    """gamma = 0 turns GAE into adv[t] = r[t] - V[t] = r[t] (V = 0), r = its own index: any element read out
    of a stage the DMA is already refilling (or not yet filled) shows up as a wrong integer.  Ring depth 2,
    many tiles per CTA, repeated with warm caches (the refill then lands fastest).  Guards the proxy fence
    in release_stage()."""
    if layout == "time_major":
        T, B = 16384, 64
        r = (torch.arange(T, device="cuda", dtype=torch.float32)[:, None] * 64
             + torch.arange(B, device="cuda", dtype=torch.float32)[None, :]).contiguous()
        opts = dict(tm_kernel=6, strip_stages=2)
    else:
        B, T = 16, 16384
        r = (torch.arange(T, device="cuda", dtype=torch.float32)[None, :]
             + 100000.0 * torch.arange(B, device="cuda", dtype=torch.float32)[:, None]).contiguous()
        opts = dict(bm_kernel=2, row_stages=2)
    z = torch.zeros_like(r)
    with _options(**opts):
        for _ in range(6):
            adv, ret = M.gae(r, z, 0.0, 0.0, next_value=z if layout == "batch_major" else None, layout=layout,
                             mode="parallel")
            assert torch.equal(adv, r)
            out = M.discounted_reward(r, 0.0, layout=layout, mode="parallel")
            assert torch.equal(out, r)


def test_tensor_array_stages_a_trajectory_for_the_scan():
    """TensorArray (utils/tensor_array.py:25-190): the docstring example (:41-58), fixed-size and dynamic
    semantics, and the A2C use (a2c.py:116-164): write(t, ...) per step, stack(), DiscountedReturn."""
    ta = M.TensorArray(np.int64, ())
    ta.write(0, 1)
    ta.write(1, 2)
    assert int(ta.read(1).cpu()) == 2
    assert ta.stack().cpu().tolist() == [1, 2] and ta.size() == 2
    ta.clear()
    ta.write(0, 3)
    assert int(ta.read(0).cpu()) == 3 and ta.size() == 1
    for i in range(1, 100):  # grows past its first allocation, keeps what it holds
        ta.write(i, i * 7)
    assert ta.stack().cpu().tolist() == [3] + [i * 7 for i in range(1, 100)]
    ta.close()
    with pytest.raises(RuntimeError):
        ta.read(0)

    T, B = 64, 8
    rng = np.random.default_rng(9)
    r = rng.normal(size=(T, B)).astype(np.float32)
    d = rng.random((T, B)) < 0.1
    rewards = M.TensorArray(np.float32, (B,), dynamic_size=False, size=T)
    dones = M.TensorArray(np.bool_, (B,), dynamic_size=False, size=T)
    for t in range(T):
        rewards.write(t, dev(r[t]) if t % 2 else r[t])  # device and host values
        dones.write(t, dev(d[t]))
    with pytest.raises(IndexError):
        rewards.write(T, r[0])
    with pytest.raises(ValueError):
        rewards.write(0, np.zeros(B + 1, np.float32))
    assert tuple(rewards.stack().shape) == (T, B)
    assert np.array_equal(rewards.stack().cpu().numpy(), r) and np.array_equal(dones.view().cpu().numpy(), d)
    lv = rng.normal(size=(B,)).astype(np.float32)
    want = O.discounted_return(r, d, lv, 0.99)
    exact(M.DiscountedReturn(0.99, mode="sequential")(rewards.view(), dones.view(), dev(lv)), want)
    close(M.DiscountedReturn(0.99)(rewards.stack(), dones.stack(), dev(lv)), want)
    rewards.clear()
    assert rewards.size() == 0 and not rewards.stack().cpu().numpy().any()
    with pytest.raises(TypeError):
        M.TensorArray("complex64", ())
    with pytest.raises(ValueError):
        M.TensorArray(np.float32, (), size=-1)


@pytest.mark.parametrize("chunks", [1, 3, 16])
def test_host_scans_pipelined_in_pieces(chunks):
    """*_host entry points cut the batch axis into pieces (copy-in / scan / copy-out overlapped): ragged
    last piece, trailing dims, both layouts, every optional array; results identical to the device path"""
    with _options(host_scan_chunks=chunks):
        rng = np.random.default_rng(chunks)
        for T, B, E in [(50, 100, 1), (33, 70, 3), (7, 1, 1), (128, 333, 1)]:
            shape = (T, B, E) if E > 1 else (T, B)
            r = rng.normal(size=shape).astype(np.float32)
            d = rng.random((T, B)) < 0.05
            lv = rng.normal(size=shape[1:]).astype(np.float32)
            got = M.DiscountedReturn(0.99)(r, d, lv)   # numpy in -> *_host path, numpy out
            assert isinstance(got, np.ndarray)
            close(got, O.discounted_return(r, d, lv, 0.99))
            exact(M.DiscountedReturn(0.99, mode="sequential")(r, d, lv), O.discounted_return(r, d, lv, 0.99))
            if E == 1:
                v = rng.normal(size=(T, B)).astype(np.float32)
                nv = rng.normal(size=(T, B)).astype(np.float32)
                adv_r, ret_r = O.gae(r, v, None, lv, d, 0.99, 0.95, layout=0)
                adv, ret = M.gae(r, v, 0.99, 0.95, last_value=lv, done=d, mode="sequential")
                exact(adv, adv_r)
                exact(ret, ret_r)
                # batch-major: the same arrays read as [T envs, B steps]
                adv_r, ret_r = O.gae(r, v, nv, None, d, 0.99, 0.95, layout=1)
                adv, ret = M.gae(r, v, 0.99, 0.95, next_value=nv, done=d, layout="batch_major", mode="sequential")
                exact(adv, adv_r)
                exact(ret, ret_r)
                adv, ret = M.gae(r, v, 0.99, 0.95, next_value=nv, layout="batch_major", want_returns=False)
                assert ret is None
                close(adv, O.gae(r, v, nv, None, None, 0.99, 0.95, layout=1)[0])


# ---- round 2: fixtures run by the reference's own statements, fused neighbours of the scan, graphs ---------
def test_mappo_gae_golden(golden_dir):
    """MAPPO's GAE block (mappo.py:478-500) as run by the reference's own statements: the sequential kernel gives
    the reference's discounted_r bit for bit, the parallel kernels within the scans' tolerance"""
    g = np.load(os.path.join(golden_dir, "mappo_gae_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        reward, value, mask = g[f"{n}/reward"][..., 0], g[f"{n}/value"][..., 0], g[f"{n}/mask"][..., 0]
        args = (dev(reward[1:]), dev(value[1:]), float(g[f"{n}/gamma"]), float(g[f"{n}/lam"]))
        kw = dict(last_value=dev(g[f"{n}/last_value"][..., 0]), done=dev(1.0 - mask[1:]))
        adv, ret = M.gae(*args, **kw, mode="sequential")
        exact(ret, g[f"{n}/discounted_r"][1:, :, 0])
        for mode in ("auto", "parallel"):
            adv, ret = M.gae(*args, **kw, mode=mode)
            close(ret, g[f"{n}/discounted_r"][1:, :, 0])
            close(adv, g[f"{n}/advantage"][..., 0])


def test_coma_td_lambda_golden(golden_dir):
    """COMA's build_td_lambda_targets (coma.py:489-501) as run by the reference's own method: bit-exact"""
    g = np.load(os.path.join(golden_dir, "coma_td_lambda_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        got = M.td_lambda_targets(dev(g[f"{n}/rewards"]), dev(g[f"{n}/terminated"]), dev(g[f"{n}/mask"]),
                                  dev(g[f"{n}/target_qs"]), float(g[f"{n}/gamma"]), float(g[f"{n}/td_lambda"]),
                                  max_t=int(g[f"{n}/max_t"]))
        exact(got, g[f"{n}/ret"])
    # rewards / mask / terminated with a full agent axis, and against the oracle at a size the fixtures do not have
    rng = np.random.default_rng(3)
    B, T, A = 40, 200, 6
    r, te = rng.normal(size=(B, T, A)).astype(np.float32), (rng.random((B, T, A)) < 0.03).astype(np.float32)
    m, q = (rng.random((B, T, A)) > 0.1).astype(np.float32), rng.normal(size=(B, T + 1, A)).astype(np.float32)
    exact(M.td_lambda_targets(dev(r), dev(te), dev(m), dev(q), 0.99, 0.8), O.td_lambda_targets(r, te, m, q, 0.99, 0.8))


@pytest.mark.parametrize("T,B", [(1, 4), (5, 3), (20, 16), (100, 8), (64, 2048), (300, 4112), (2048, 256), (512, 40000)])
def test_vtrace_fused_against_oracle(T, B):
    """mrl_vtrace (exp, clips, deltas, recurrence, vs, pg advantages in the library) against the oracle's restatement
    of vtrace.py:79-105 in all three modes; thresholds that clip, thresholds that do not, and no masks"""
    rng = np.random.default_rng(T * 7 + B)
    lr = (rng.normal(size=(T, B)) * 0.7).astype(np.float32)
    disc = (0.99 * (rng.random((T, B)) > 0.02)).astype(np.float32)
    r, v = rng.normal(size=(T, B)).astype(np.float32), rng.normal(size=(T, B)).astype(np.float32)
    boot = rng.normal(size=(B,)).astype(np.float32)
    masks = (rng.random((T, B)) > 0.1).astype(np.float32)
    for clips, mk in (((1.0, 1.0, 1.0), masks), ((1.5, 2.0, 0.9), masks), ((None, None, None), None)):
        want_vs, want_pg = O.vtrace(lr, disc, r, v, boot, *clips, mk)
        for mode in ("auto", "sequential", "parallel"):
            vs, pg = M.get_vs_and_advantages(dev(lr), dev(disc), dev(r), dev(v), dev(boot), *clips,
                                             None if mk is None else dev(mk), mode=mode)
            close(vs, want_vs)
            close(pg, want_pg)


@pytest.mark.parametrize("layout,T,B,with_done", [("time_major", 2048, 4096, True), ("time_major", 300, 4112, False),
                                                  ("time_major", 1000, 96, True), ("batch_major", 2048, 512, True),
                                                  ("batch_major", 1024, 300, False), ("batch_major", 130, 17, True),
                                                  ("time_major", 5, 3, True)])
def test_gae_with_fused_normalisation(layout, T, B, with_done):
    """gae(normalize=True): moments accumulated by the scan kernel while it stores (TMA strips / rows) or in a
    separate pass (the other kernels) -- either way equal to normalising the un-normalised result in float64, and
    the returns are untouched (ppo.py:353-368)"""
    rng = np.random.default_rng(T + B)
    shape = (T, B) if layout == "time_major" else (B, T)
    r, v = rng.normal(size=shape).astype(np.float32), rng.normal(size=shape).astype(np.float32)
    lv = rng.normal(size=(B,)).astype(np.float32)
    d = (rng.random(shape) < 0.01) if with_done else None
    kw = dict(last_value=dev(lv), done=None if d is None else dev(d), layout=layout)
    adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, **kw)
    nadv, nret = M.gae(dev(r), dev(v), 0.99, 0.95, **kw, normalize=True)
    # the TMA strip / row kernels chain their tiles in a fixed order: bit-reproducible.  The look-back kernels of
    # narrow or unaligned shapes compose whatever aggregates have been published when a tile looks back, so their
    # last bits depend on timing: held to the scans' tolerance instead
    fixed_order = (layout == "time_major" and B >= 2048 and B % 16 == 0) or (layout == "batch_major" and T >= 512 and T % 16 == 0) \
        or T * B < 64
    (exact if fixed_order else close)(nret, ret.cpu().numpy())
    a = adv.cpu().numpy()
    a64 = a.astype(np.float64)
    want = (a - np.float32(a64.mean())) * (np.float32(1.0) / (np.sqrt(np.float32(a64.var())) + np.float32(1e-8)))
    assert np.allclose(nadv.cpu().numpy(), want, rtol=3e-6, atol=3e-6)
    again, _ = M.gae(dev(r), dev(v), 0.99, 0.95, **kw, normalize=True)
    if fixed_order:
        exact(again, nadv.cpu().numpy())  # fixed accumulation order of the moments
    # A2C: returns + normalisation (a2c.py:189-192)
    if layout == "time_major" and d is not None:
        out = M.DiscountedReturn(0.99)(dev(r), dev(d), dev(lv)).cpu().numpy()
        o64 = out.astype(np.float64)
        want = (out - np.float32(o64.mean())) * (np.float32(1.0) / (np.sqrt(np.float32(o64.var())) + np.float32(1e-8)))
        got = M.discounted_return_normalized(dev(r), dev(d), dev(lv), 0.99).cpu().numpy()
        assert np.allclose(got, want, rtol=3e-6, atol=3e-6)


@pytest.mark.parametrize("T,B", [(1000, 96), (2048, 128), (2048, 4096), (513, 36)])
def test_scans_replay_in_a_cuda_graph(T, B):
    """the look-back kernels kept their epoch / ticket base on the host and could not be captured.
    They now advance both in the workspace (last CTA), so a captured mrl_gae / mrl_discounted_return replays:
    look-back shapes (B < 2048 or unaligned) and TMA-strip shapes, fresh inputs on every replay, oracle each time."""
    rng = np.random.default_rng(T * B)
    r_d = torch.zeros((T, B), device="cuda")
    v_d = torch.zeros((T + 1, B), device="cuda")
    d_d = torch.zeros((T, B), dtype=torch.bool, device="cuda")

    def fresh():
        r = rng.normal(size=(T, B)).astype(np.float32)
        v = rng.normal(size=(T + 1, B)).astype(np.float32)
        d = rng.random((T, B)) < 0.01
        r_d.copy_(torch.from_numpy(r)), v_d.copy_(torch.from_numpy(v)), d_d.copy_(torch.from_numpy(d))
        return O.gae(r, v[:T], None, v[T], d, 0.99, 0.95, layout=0), O.discounted_return(r, d, v[T], 0.99)

    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        want = fresh()
        adv, ret = M.gae(r_d, v_d[:T], 0.99, 0.95, last_value=v_d[T], done=d_d)  # eager: allocates the workspaces
        dr = M.DiscountedReturn(0.99)(r_d, d_d, v_d[T])
        M.gae(r_d, v_d[:T], 0.99, 0.95, last_value=v_d[T], done=d_d, normalize=True)
        close(adv, want[0][0]), close(dr, want[1])
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            adv, ret = M.gae(r_d, v_d[:T], 0.99, 0.95, last_value=v_d[T], done=d_d)
            dr = M.DiscountedReturn(0.99)(r_d, d_d, v_d[T])
            nadv, _ = M.gae(r_d, v_d[:T], 0.99, 0.95, last_value=v_d[T], done=d_d, normalize=True)
        for _ in range(4):
            want = fresh()
            graph.replay()
            torch.cuda.synchronize()
            close(adv, want[0][0]), close(ret, want[0][1]), close(dr, want[1])
            a = want[0][0].astype(np.float64)
            assert np.allclose(nadv.cpu().numpy(), (want[0][0] - a.mean()) / (a.std() + 1e-8), rtol=1e-4, atol=1e-4)
        # and eager launches still work on the same stream after the replays (the state lives in the workspace)
        want = fresh()
        adv2, _ = M.gae(r_d, v_d[:T], 0.99, 0.95, last_value=v_d[T], done=d_d)
        close(adv2, want[0][0])


def test_scans_on_two_streams_at_once():
    """the look-back workspace, epoch and ticket base were process-global, so two scans in flight on
    different streams corrupted each other (or hung).  Workspaces are per (device, stream) now: interleave many
    launches of a look-back shape on two streams without synchronising in between."""
    rng = np.random.default_rng(99)
    T, B = 1500, 160  # look-back kernels (B < 2048)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    data = []
    for k in range(2):
        r, v = rng.normal(size=(T, B)).astype(np.float32), rng.normal(size=(T + 1, B)).astype(np.float32)
        d = rng.random((T, B)) < 0.01
        data.append((dev(r), dev(v), dev(d), O.gae(r, v[:T], None, v[T], d, 0.99, 0.95, layout=0)))
    torch.cuda.synchronize()
    outs = [[], []]
    for it in range(30):
        for k in range(2):
            with torch.cuda.stream(streams[k]):
                r, v, d, _ = data[k]
                outs[k].append(M.gae(r, v[:T], 0.99, 0.95, last_value=v[T], done=d))
    torch.cuda.synchronize()
    for k in range(2):
        for adv, ret in outs[k]:
            close(adv, data[k][3][0]), close(ret, data[k][3][1])
