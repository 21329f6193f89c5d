"""GPU parity of UniformReplayBuffer (BufferAppend / BufferGetItem / BufferSample) against the CPU
oracle: cursor state, column bytes, logical indexing, drawn slots and gathered rows, all bit-exact."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import oracle as O  # noqa: E402

import mindrl_b200 as M  # noqa: E402
from mindrl_b200 import _ffi  # noqa: E402

SHAPES = [(4,), (1,), (1,), (4,)]
TYPES = [np.float32, np.int32, np.float32, np.float32]


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def rows(rng, n):
    return [rng.normal(size=(n, 4)).astype(np.float32), rng.integers(0, 2, size=(n, 1)).astype(np.int32),
            rng.normal(size=(n, 1)).astype(np.float32), rng.normal(size=(n, 4)).astype(np.float32)]


def same_columns(gpu, orc):
    for g, o in zip(gpu.buffer, orc.columns()):
        assert np.array_equal(g.cpu().numpy().view(np.uint8), o.view(np.uint8))


@pytest.mark.parametrize("carrier", ["device", "host"])
def test_append_get_item_sequence(carrier):
    rng = np.random.default_rng(0)
    cap = 100
    g = M.UniformReplayBuffer(8, cap, SHAPES, TYPES, seed=3)
    o = O.UniformBufferOracle(8, cap, SHAPES, TYPES, seed=3)
    assert g.count == 0 and g.head == 0 and not g.full()[0] and g.size() == 0
    for n in (1, 1, 7, 50, 41, 1, 13, 99, 100, 250, 1, 3):
        batch = rows(rng, n)
        if n == 1:  # single transition without the batch dim, as trainers pass it
            batch = [b[0] for b in batch]
        ret = g.insert([dev(b) for b in batch] if carrier == "device" else batch)
        assert ret is g.buffer
        o.insert(batch)
        assert (g.count, g.head) == o.state()
        same_columns(g, o)
        for idx in (0, 1, g.count - 1, -1, g.count // 2):
            if -g.count <= idx < g.count:
                got = g.get_item(idx)
                want = o.get_item(idx)
                for a, b in zip(got, want):
                    assert np.array_equal(a.cpu().numpy(), b)
    assert g.full()[0] and int(g.size()) == cap
    with pytest.raises(IndexError):
        g.get_item(cap)
    assert g.reset() is True
    assert g.count == 0 and g.head == o.state()[1]


def test_sample_draws_and_rows_match_oracle():
    rng = np.random.default_rng(1)
    cap = 1000
    g = M.UniformReplayBuffer(64, cap, SHAPES, TYPES, seed=11)
    o = O.UniformBufferOracle(64, cap, SHAPES, TYPES, seed=11)
    with pytest.raises(RuntimeError):
        g.sample()
    batch = rows(rng, 600)
    g.insert([dev(b) for b in batch])
    o.insert(batch)
    for _ in range(5):
        got = g.sample()
        slots, want = o.sample()
        for a, b in zip(got, want):
            assert tuple(a.shape) == b.shape and np.array_equal(a.cpu().numpy(), b)
    # explicit slots, host carrier
    gh = M.UniformReplayBuffer(64, cap, SHAPES, TYPES, seed=11, carrier="numpy")
    gh.insert(batch)
    for _ in range(2):
        o2 = O.UniformBufferOracle(64, cap, SHAPES, TYPES, seed=11)
        o2.insert(batch)
        got = gh.sample()
        _, want = o2.sample()
        assert all(isinstance(a, np.ndarray) for a in got)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        break
    slots = rng.integers(0, 600, size=333)
    got = g.gather(dev(slots))
    want = o.gather(slots)
    for a, b in zip(got, want):
        assert np.array_equal(a.cpu().numpy(), b)


@pytest.mark.parametrize("path", ["bulk", "ldg"])
def test_wide_rows_gather_paths(path):
    """Atari-shaped rows (84*84*4 bytes) through the TMA ring and through the register path, plus
    odd row sizes; rows are the deterministic fill so every byte is checked."""
    cap, n = 2000, 257
    _ffi.set_option("gather_path", 2 if path == "bulk" else 1)
    for row_bytes in (28224, 4096, 28224 + 16, 1000, 6, 3):
        rb = np.asarray([row_bytes, 8], dtype=np.int64)
        h = C.c_int64(-1)
        _ffi.check(_ffi.lib().mrl_urb_create(cap, 2, rb.ctypes.data, None, 0, C.byref(h)))
        cols = (C.c_void_p * 2)()
        _ffi.check(_ffi.lib().mrl_urb_columns(h.value, cols))
        _ffi.check(_ffi.lib().mrl_fill_rows(cols[0], cap, row_bytes, 77, 0))
        _ffi.check(_ffi.lib().mrl_fill_rows(cols[1], cap, 8, 78, 0))
        ref0, ref1 = O.fill_rows(cap, row_bytes, 77), O.fill_rows(cap, 8, 78)
        slots = np.random.default_rng(row_bytes).integers(0, cap, size=n)
        slots[:3] = [0, cap - 1, cap - 1]
        ds = dev(slots)
        out0 = torch.empty((n, row_bytes), dtype=torch.uint8, device="cuda")
        out1 = torch.empty((n, 8), dtype=torch.uint8, device="cuda")
        tab = _ffi.ptr_table([out0.data_ptr(), out1.data_ptr()])
        _ffi.check(_ffi.lib().mrl_urb_gather(h.value, ds.data_ptr(), n, tab, 0))
        torch.cuda.synchronize()
        assert np.array_equal(out0.cpu().numpy(), ref0[slots])
        assert np.array_equal(out1.cpu().numpy(), ref1[slots])
        if row_bytes >= 8:
            assert np.array_equal(out0.cpu().numpy()[:, :8].copy().view(np.int64)[:, 0], slots)
        _ffi.check(_ffi.lib().mrl_urb_destroy(h.value))
    _ffi.set_option("gather_path", 0)


def test_bulk_append_wide_rows_and_wrap():
    cap = 300
    shapes, types = [(84, 84, 4), (1,)], [np.uint8, np.int32]
    g = M.UniformReplayBuffer(16, cap, shapes, types)
    o = O.UniformBufferOracle(16, cap, shapes, types)
    rng = np.random.default_rng(2)
    for n in (200, 170, 5):
        obs = rng.integers(0, 256, size=(n, 84, 84, 4), dtype=np.uint8)
        act = rng.integers(0, 18, size=(n, 1)).astype(np.int32)
        g.insert([dev(obs), dev(act)])
        o.insert([obs, act])
        assert (g.count, g.head) == o.state()
    same_columns(g, o)
    got, want = g.sample(), None
    assert tuple(got[0].shape) == (16, 84, 84, 4)


def test_launch_counter_moves():
    before = _ffi.launch_count()
    g = M.UniformReplayBuffer(4, 10, [(1,)], [np.float32])
    g.insert([dev(np.ones((5, 1), np.float32))])
    g.sample()
    assert _ffi.launch_count() >= before + 2


def test_n_step_buffer_and_checkpoint():
    """NStepBuffer (utils/n_step_buffer.py:54-103) over get_item, and state_dict round trip"""
    shapes, types = [(4,), (1,)], [np.float32, np.int32]
    nb = M.NStepBuffer(shapes, types, td_step=3, buffer_size=50)
    check, _ = nb.get_data()
    assert not check
    for i in range(7):
        nb.push([dev(np.full(4, i, np.float32)), dev(np.array([i], np.int32))])
    check, cols = nb.get_data()
    assert check
    assert np.array_equal(cols[1].cpu().numpy()[:3, 0], [4, 5, 6])
    assert np.array_equal(cols[0].cpu().numpy()[:3, 0], [4.0, 5.0, 6.0])
    assert nb.clear() is True and nb.buffer.count == 0
    g = M.UniformReplayBuffer(4, 20, shapes, types, seed=2)
    rng = np.random.default_rng(3)
    g.insert([dev(rng.normal(size=(27, 4)).astype(np.float32)), dev(np.arange(27, dtype=np.int32)[:, None])])
    sd = g.state_dict()
    h = M.UniformReplayBuffer(4, 20, shapes, types, seed=2)
    h.load_state_dict(sd)
    assert (h.count, h.head) == (g.count, g.head)
    for a, b in zip(g.buffer, h.buffer):
        assert torch.equal(a, b)
    for i in (0, 5, -1):
        for a, b in zip(g.get_item(i), h.get_item(i)):
            assert torch.equal(a, b)


def test_reservoir_buffer_matches_oracle_and_reference_test():
    """ReservoirReplayBuffer against its numpy oracle (same seeds -> same keep/replace decisions -> identical
    columns and samples), plus the expectations of the reference's (skipped) test
    tests/st/test_reservoir_replay_buffer.py:34-56."""
    import ctypes as C
    from mindrl_b200 import _ffi
    cap, batch = 100, 256
    shapes, dtypes = [(17,), (6,)], [np.float32, np.int32]
    g = M.ReservoirReplayBuffer(cap, batch, shapes, dtypes, seed0=0, seed1=42)
    o = O.ReservoirBufferOracle(cap, batch, shapes, dtypes, seed0=0, seed1=42)
    for i in range(100):
        g.push(dev(np.ones(17, np.float32) * i), dev(np.ones(6, np.int32) * i))
        o.push(np.ones(17, np.float32) * i, np.ones(6, np.int32) * i)
    states, actions = g.sample()
    states, actions = states.cpu().numpy(), actions.cpu().numpy()
    assert states.shape == (100, 17) and actions.shape == (100, 6)  # "variable-length": min(batch, size)
    assert np.allclose(states, states[:, 0:1]) and np.allclose(actions, actions[:, 0:1])
    assert np.allclose(states[:, 0], actions[:, 0])
    slots, want = o.sample()
    assert np.array_equal(g.last_slots, slots)
    assert np.array_equal(states, want[0]) and np.array_equal(actions, want[1])
    # past capacity: host and device inputs mixed, drops and replacements
    for i in range(100, 1500):
        s, a = np.ones(17, np.float32) * i, np.ones(6, np.int32) * i
        if i % 3:
            g.push(dev(s), dev(a))
        else:
            g.push(s, a)
        o.push(s, a)
    ptrs = (C.c_void_p * 2)()
    _ffi.check(_ffi.lib().mrl_urb_columns(g._handle, ptrs))
    got = np.empty((cap, 17), np.float32)
    _ffi.check(_ffi.lib().mrl_memcpy_d2h(got.ctypes.data, ptrs[0], got.nbytes, 0))
    _ffi.check(_ffi.lib().mrl_stream_sync(0))
    assert np.array_equal(got, o.cols[0])
    assert len(np.unique(got[:, 0])) == cap and got.max() > 100  # replaced rows, no duplicates
    for _ in range(3):
        states, actions = g.sample()
        slots, want = o.sample()
        assert np.array_equal(states.cpu().numpy(), want[0]) and np.array_equal(actions.cpu().numpy(), want[1])
    with pytest.raises(ValueError):
        g.push(dev(np.ones((2, 17), np.float32)), dev(np.ones((2, 6), np.int32)))
    assert g.destroy()[0] > 0 and g.destroy()[0] == -1


@pytest.mark.parametrize("A,B,inner,dtype", [(1000, 30, (17,), np.float32), (1000, 30, (1,), np.float32),
                                             (1000, 30, (), np.float32), (33, 65, (6,), np.int32), (7, 5, (3,), np.uint8),
                                             (64, 64, (), np.float64), (100, 31, (), np.bool_), (50, 20, (5,), np.float16),
                                             (2048, 4096, (), np.float32), (1, 9, (4,), np.int64)])
def test_get_replay_buffer_elements_transpose(A, B, inner, dtype):
    """MSRL.get_replay_buffer_elements(transpose=True, shape=(1, 0, 2)) (core/msrl.py:535-557): bit-exact
    swap of the leading axes for every row width the kernels dispatch on"""
    rng = np.random.default_rng(A * 31 + B)
    x = (rng.integers(0, 200, size=(A, B) + inner)).astype(dtype)
    perm = (1, 0) + tuple(range(2, 2 + len(inner)))
    (got,) = M.get_replay_buffer_elements((dev(x),), transpose=True, shape=perm)
    assert tuple(got.shape) == (B, A) + inner
    assert np.array_equal(got.cpu().numpy(), np.transpose(x, perm))
    (same,) = M.get_replay_buffer_elements((dev(x),))
    assert np.array_equal(same.cpu().numpy(), x)
    if inner:
        with pytest.raises(ValueError):
            M.get_replay_buffer_elements((dev(x),), transpose=True, shape=(0, 2, 1)[:2 + len(inner)] if len(inner) == 1 else (1, 0))


def test_buffer_elements_feed_the_scans_without_a_transpose():
    """the PPO path (ppo_trainer.py:70 -> ppo.py:320-350): columns [T, env, 1] are transposed to [env, T, 1] and
    scanned batch-major; scanning the untransposed column time-major gives the same numbers, transposed"""
    T, env = 1000, 32
    ub = M.UniformReplayBuffer(1, T, [(env, 1), (env, 1)], [np.float32, np.float32])
    rng = np.random.default_rng(5)
    r = rng.normal(size=(T, env, 1)).astype(np.float32)
    v = rng.normal(size=(T, env, 1)).astype(np.float32)
    ub.insert([dev(r), dev(v)])
    rt, vt = ub.get_replay_buffer_elements(transpose=True, shape=(1, 0, 2))
    assert tuple(rt.shape) == (env, T, 1)
    want = O.discounted_return(r[:, :, 0].T.copy(), None, None, 0.99, layout=1)
    a = M.discounted_reward(rt.reshape(env, T), 0.99, layout="batch_major").cpu().numpy()
    b = M.discounted_reward(ub.buffer[0].reshape(T, env), 0.99, layout="time_major").cpu().numpy()
    tol = 1e-5 * max(1.0, float(np.abs(want).max()))
    assert np.abs(a - want).max() <= tol and np.abs(b.T - want).max() <= tol


def test_large_host_batches_go_straight_into_the_ring():
    """numpy batches of >= 256 KB are DMA'd from the caller's arrays into the ring rows (no staging copy, no
    kernel): same columns and cursors as the oracle through a fill, a wrap, an over-long batch, the emplace of
    the reservoir buffer, and the priority buffer's push; the arrays may be overwritten right after the call"""
    cap = 300
    shapes, types = [(84, 84, 4), (1,)], [np.uint8, np.int32]
    g = M.UniformReplayBuffer(16, cap, shapes, types, carrier="numpy")
    o = O.UniformBufferOracle(16, cap, shapes, types)
    rng = np.random.default_rng(12)
    launches = _ffi.launch_count()
    for n in (200, 170, 40, 650):
        obs = rng.integers(0, 256, size=(n, 84, 84, 4), dtype=np.uint8)
        act = rng.integers(0, 18, size=(n, 1)).astype(np.int32)
        g.insert([obs, act])
        o.insert([obs.copy(), act.copy()])
        obs[:] = 0  # the call has finished with the arrays
        act[:] = 0
        assert (g.count, g.head) == o.state()
    assert _ffi.launch_count() == launches  # copies only
    same_columns(g, o)
    # emplace at a physical slot (ReservoirReplayBufferPush)
    L = _ffi.lib()
    obs = rng.integers(0, 256, size=(20, 84, 84, 4), dtype=np.uint8)
    act = rng.integers(0, 18, size=(20, 1)).astype(np.int32)
    _ffi.check(L.mrl_urb_write_at_host(g._handle, 33, 20, _ffi.ptr_table([obs.ctypes.data, act.ctypes.data]), 0))
    cols = o.columns()
    cols[0][33:53], cols[1][33:53] = obs, act
    same_columns(g, o)
    # priority buffer: rows by DMA, then leaves and tree as always
    pg = M.PriorityReplayBuffer(0.6, cap, 16, shapes, types, 1, 2, carrier="numpy")
    po = O.PriorityBufferOracle(0.6, cap, 16, shapes, types, 1, 2)
    for n in (200, 170):
        obs = rng.integers(0, 256, size=(n, 84, 84, 4), dtype=np.uint8)
        act = rng.integers(0, 18, size=(n, 1)).astype(np.int32)
        pg.insert(obs, act)
        po.insert_batch([obs.copy(), act.copy()])
        obs[:] = 0
    gs, gm = pg.tree()
    os_, om = po.tree()
    assert np.array_equal(gs[1:], os_[1:]) and np.array_equal(gm[1:], om[1:]) and pg.info() == po.info()
    u = rng.random(16).astype(np.float32)
    got, want = pg.sample(0.4, uniforms=u), po.sample(0.4, uniforms=u)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2]) and np.array_equal(got[3], want[3])


@pytest.mark.parametrize("carrier", ["device", "numpy"])
def test_c1_dqn_shape_loop_bit_exact(carrier):
    """BASELINE config 1 at its own size: capacity 100 000, batch 64, the DQN loop's {insert 1 row, sample 64} --
    through the fill phase (count < capacity), the moment the ring gets full, and the wrap-around after it.
    Cursors after every step; drawn rows bit-exact; whole columns at the three phase boundaries."""
    rng = np.random.default_rng(21)
    cap, batch = 100_000, 64
    g = M.UniformReplayBuffer(batch, cap, SHAPES, TYPES, seed=5, carrier=carrier)
    o = O.UniformBufferOracle(batch, cap, SHAPES, TYPES, seed=5)
    to = (lambda a: dev(a)) if carrier == "device" else (lambda a: a)
    pre = rows(rng, cap - 150)
    g.insert([to(c) for c in pre])
    o.insert(pre)
    steps = rows(rng, 400)
    for i in range(400):
        g.insert([to(c[i]) for c in steps])
        o.insert([c[i] for c in steps])
        assert (int(g.count), int(g.head)) == o.state()
        got = g.sample()
        _, want = o.sample()
        for a, b in zip(got, want):
            assert np.array_equal(a.cpu().numpy() if hasattr(a, "cpu") else a, b)
        if i in (0, 149, 150, 399):
            same_columns(g, o) if carrier == "device" else None
    assert bool(g.full()[0]) and o.state() == (cap, 250)
    for index in (0, -1, cap - 1, 12345):
        for a, b in zip(g.get_item(index), o.get_item(index)):
            assert np.array_equal((a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)).reshape(-1), np.asarray(b).reshape(-1))
