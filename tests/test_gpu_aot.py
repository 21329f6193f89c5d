"""The MindSpore "aot" custom-operator entry points (include/mindrl_b200_aot.h) called the way MindSpore's
AOT kernel mod calls them -- fabricated params / ndims / shapes / dtypes / AotExtra -- and checked against the
CPU oracle.  These are the symbols ops.Custom("libmindrl_b200.so:<Func>", ..., "aot") would bind
(reference ABI: mindspore_rl/utils/mcts/cpu/cpu_mcts_ops.cc:38-53)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import oracle as O  # noqa: E402
from aot_host import AotOp, t  # noqa: E402

SHAPES, TYPES = [(4,), (1,), (1,), (4,)], [np.float32, np.int32, np.float32, np.float32]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def rows_of(rng, n):
    return [rng.normal(size=(n, 4)).astype(np.float32), rng.integers(0, 5, (n, 1)).astype(np.int32),
            rng.normal(size=(n, 1)).astype(np.float32), rng.normal(size=(n, 4)).astype(np.float32)]


def stream():
    return torch.cuda.current_stream().cuda_stream


def test_uniform_buffer_ops_against_oracle():
    """BufferAppend / BufferGetItem / BufferSample on caller-owned columns with DEVICE cursors: one row, batches,
    wrap-around, an over-long batch, negative get_item indices; columns, count and head bit-exact"""
    rng = np.random.default_rng(5)
    cap, batch = 257, 32
    buf = [torch.zeros((cap,) + s, dtype=getattr(torch, np.dtype(d).name), device="cuda") for s, d in zip(SHAPES, TYPES)]
    count = torch.zeros(1, dtype=torch.int32, device="cuda")
    head = torch.zeros(1, dtype=torch.int32, device="cuda")
    count_out = torch.zeros(1, dtype=torch.int32, device="cuda")
    o = O.UniformBufferOracle(batch, cap, SHAPES, TYPES, seed=11)
    append, get, sample = AotOp("BufferAppend"), AotOp("BufferGetItem"), AotOp("BufferSample", seed=11)
    for n in [1, 1, 40, 200, 1, 300, 5, 700, 1]:
        rows = rows_of(rng, n)
        exp = [dev(c[0] if n == 1 else c) for c in rows]
        assert append([t(b) for b in buf] + [t(e) for e in exp] + [t(count), t(head), t(count_out)], stream()) == 0
        o.insert([c[0] for c in rows] if n == 1 else rows)
        c_ref, h_ref = o.state()
        assert (int(count.item()), int(head.item()), int(count_out.item())) == (c_ref, h_ref, c_ref)
        for b, ob in zip(buf, o.columns()):
            assert np.array_equal(b.cpu().numpy(), ob)
        for index in (0, c_ref - 1, -1, -c_ref, c_ref // 2):
            idx = torch.tensor([index], dtype=torch.int64, device="cuda")
            outs = [torch.empty(s, dtype=b.dtype, device="cuda") for s, b in zip(SHAPES, buf)]
            assert get([t(b) for b in buf] + [t(count), t(head), t(idx)] + [t(x) for x in outs], stream()) == 0
            for a, b in zip(outs, o.get_item(index)):
                assert np.array_equal(a.cpu().numpy(), np.asarray(b).reshape(a.shape))
        if c_ref >= batch:
            outs = [torch.empty((batch,) + s, dtype=b.dtype, device="cuda") for s, b in zip(SHAPES, buf)]
            assert sample([t(b) for b in buf] + [t(count), t(head)] + [t(x) for x in outs], stream()) == 0
            _, want = o.sample()
            for a, b in zip(outs, want):
                assert np.array_equal(a.cpu().numpy(), b)


def test_priority_buffer_ops_against_oracle():
    """Create / Push / Sample(beta as a tensor) / Update / Destroy through the AOT signature: indices, weights,
    rows and the whole tree bit-exact against the oracle"""
    from mindrl_b200 import _ffi
    import ctypes as C
    rng = np.random.default_rng(9)
    cap, batch, alpha = 3000, 64, 0.6
    create = AotOp("PriorityReplayBufferCreate", capacity=cap, alpha=alpha, shapes=[list(s) for s in SHAPES],
                   dtypes="float32, int32,float32,float32", seed0=3, seed1=4)
    hbuf = torch.full((1,), -7, dtype=torch.int64, device="cuda")
    assert create([t(hbuf)], stream()) == 0
    handle = int(hbuf.item())
    assert handle > 0
    o = O.PriorityBufferOracle(alpha, cap, batch, SHAPES, TYPES, 3, 4)
    push, sample = AotOp("PriorityReplayBufferPush", handle=handle), AotOp("PriorityReplayBufferSample", handle=handle)
    update, destroy = AotOp("PriorityReplayBufferUpdate", handle=handle), AotOp("PriorityReplayBufferDestroy", handle=handle)
    rows = rows_of(rng, 500)
    out_h = torch.zeros(1, dtype=torch.int64, device="cuda")
    drows = [dev(c) for c in rows]  # (kept alive: t() only takes pointers)
    for i in range(500):
        assert push([t(c[i]) for c in drows] + [t(out_h)], stream()) == 0
        o.insert(*[c[i] for c in rows])
    assert int(out_h.item()) == handle
    assert push([t(drows[0][0]), t(out_h)], stream()) == 2  # one tensor for a buffer of four columns
    # the internal generator draws the uniforms on both sides (same seed, same counter)
    for step in range(6):
        beta = torch.tensor([0.4 + 0.1 * step], dtype=torch.float32, device="cuda")
        idx = torch.empty(batch, dtype=torch.int64, device="cuda")
        w = torch.empty(batch, dtype=torch.float32, device="cuda")
        outs = [torch.empty((batch,) + s, dtype=getattr(torch, np.dtype(d).name), device="cuda") for s, d in zip(SHAPES, TYPES)]
        assert sample([t(beta), t(idx), t(w)] + [t(x) for x in outs], stream()) == 0
        want = o.sample(float(np.float32(0.4 + 0.1 * step)))
        assert np.array_equal(idx.cpu().numpy(), want[0])
        assert np.array_equal(w.cpu().numpy().view(np.uint32), want[1].view(np.uint32))
        for a, b in zip(outs, want[2:]):
            assert np.array_equal(a.cpu().numpy(), b)
        pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
        pr_d = dev(pr)
        assert update([t(idx), t(pr_d), t(out_h)], stream()) == 0
        o.update_priorities(want[0], pr)
    L = _ffi.lib()
    s_ = np.empty(2 * 4096, np.float32)
    m_ = np.empty(2 * 4096, np.float32)
    _ffi.check(L.mrl_prb_tree_dump(handle, s_.ctypes.data, m_.ctypes.data, stream()))
    os_, om = o.tree()
    assert np.array_equal(s_[1:], os_[1:]) and np.array_equal(m_[1:], om[1:])
    assert destroy([t(out_h)], stream()) == 0
    assert L.mrl_prb_info(handle, None, None, None) == _ffi.MRL_ERR_HANDLE  # the handle is gone
    assert update([t(idx), t(pr_d), t(out_h)], stream()) == 2               # kErrorCode of the reference's ops


def test_discounted_return_and_gae_ops():
    """DiscountedReturn on the reference's two known-answer vectors (tests/st/test_discounted_return.py:39-53) and
    a random [T,B,E] case; GaeScan (+ normalize) against the oracle"""
    op = AotOp("DiscountedReturn", gamma=0.99)
    r = dev(np.ones((4, 1), np.float32))
    d = dev(np.array([[False], [False], [True], [False]]))
    last = dev(np.array([2.0], np.float32))
    out = torch.empty_like(r)
    assert op([t(r), t(d), t(last), t(out)], stream()) == 0
    assert np.allclose(out.cpu().numpy(), [[2.9701], [1.99], [1.0], [2.98]], rtol=1e-6)
    rng = np.random.default_rng(2)
    T, B, E = 96, 24, 3
    rw = rng.normal(size=(T, B, E)).astype(np.float32)
    dn = rng.random((T, B)) < 0.05
    lv = rng.normal(size=(B, E)).astype(np.float32)
    out = torch.empty((T, B, E), device="cuda")
    op2 = AotOp("DiscountedReturn", gamma=0.97)
    keep = [dev(rw), dev(dn), dev(lv)]
    assert op2([t(x) for x in keep] + [t(out)], stream()) == 0
    want = O.discounted_return(rw, dn, lv, 0.97)
    assert np.abs(out.cpu().numpy() - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max()))

    T, B = 128, 48
    rw = rng.normal(size=(T, B)).astype(np.float32)
    v = rng.normal(size=(T + 1, B)).astype(np.float32)
    dn = rng.random((T, B)) < 0.03
    adv_ref, ret_ref = O.gae(rw, v[:T], None, v[T], dn, 0.99, 0.95, layout=0)
    for normalize in (False, True):
        g = AotOp("GaeScan", gamma=0.99, **{"lambda": 0.95}, layout=0, normalize=normalize)
        adv, ret = torch.empty((T, B), device="cuda"), torch.empty((T, B), device="cuda")
        vd, rd, dd = dev(v), dev(rw), dev(dn)
        assert g([t(rd), t(vd[:T]), t(vd[T]), t(dd), t(adv), t(ret)], stream()) == 0
        want = adv_ref
        if normalize:
            a64 = adv_ref.astype(np.float64)
            want = ((adv_ref - np.float32(a64.mean())) * np.float32(1.0 / (np.sqrt(np.float32(a64.var())) + np.float32(1e-8))))
        assert np.abs(adv.cpu().numpy() - want).max() <= 2e-5 * max(1.0, float(np.abs(want).max()))
        assert np.abs(ret.cpu().numpy() - ret_ref).max() <= 1e-5 * max(1.0, float(np.abs(ret_ref).max()))


def test_bad_calls_return_the_reference_error_code():
    op = AotOp("BufferSample", seed=1, unique=True)
    x = torch.zeros((8, 2), device="cuda")
    c = torch.zeros(1, dtype=torch.int32, device="cuda")
    y = torch.zeros((2, 2), device="cuda")
    assert op([t(x), t(c), t(c), t(y)], stream()) == 2
    bad = AotOp("BufferAppend")
    assert bad([t(x), t(c)], stream()) == 2  # parameter count does not fit buffer/exp/count/head/out
