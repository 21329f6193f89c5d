"""GPU parity of PriorityReplayBuffer against the CPU oracle.  Bit-exact: sampled indices, every
node of the sum tree and the min tree, max_priority, gathered rows, importance weights."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import oracle as O  # noqa: E402

import mindrl_b200 as M  # noqa: E402
from mindrl_b200 import _ffi  # noqa: E402


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t)


def same_tree(g, o):
    gs, gm = g.tree()
    os_, om = o.tree()
    assert np.array_equal(gs[1:].view(np.uint32), os_[1:].view(np.uint32)), "sum tree differs"
    assert np.array_equal(gm[1:].view(np.uint32), om[1:].view(np.uint32)), "min tree differs"
    assert g.max_priority() == o.max_priority()
    assert g.info() == o.info()


def same_sample(got, want):
    assert np.array_equal(host(got[0]), want[0]), "indices differ"
    assert np.array_equal(host(got[1]).view(np.uint32), want[1].view(np.uint32)), "weights differ"
    for a, b in zip(got[2:], want[2:]):
        assert np.array_equal(host(a), b), "rows differ"


def test_reference_behaviour():
    # tests/st/test_priority_replay_buffer.py:36-66 with insert() instead of the missing push()
    capacity, batch = 200, 32
    prb = M.PriorityReplayBuffer(1.0, capacity, batch, ((17,), (6,)), (np.float32, np.int32), seed0=0, seed1=42)
    for i in range(100):
        h = prb.insert(dev(np.ones(17, np.float32) * i), dev(np.ones(6, np.int32) * i))
    assert h.shape == (1,) and h.dtype == np.int64
    indices, weights, states, actions = prb.sample(1.0)
    idx = host(indices)
    assert idx.dtype == np.int64 and np.all(idx < 100)
    assert np.allclose(host(states), np.broadcast_to(idx.reshape(-1, 1), states.shape))
    assert np.allclose(host(actions), np.broadcast_to(idx.reshape(-1, 1), actions.shape))
    prb.update_priorities(indices, dev(np.ones(batch, np.float32) * 1e-7))
    indices_new, _, states_new, actions_new = prb.sample(1.0)
    idx2 = host(indices_new)
    assert np.all(idx2 < 100) and np.all(idx != idx2)
    assert np.allclose(host(states_new), np.broadcast_to(idx2.reshape(-1, 1), states_new.shape))
    assert prb.full() == False  # noqa: E712  (constant, :122-124)
    with pytest.raises(ValueError):
        prb.reset()
    assert prb.destroy().shape == (1,)


@pytest.mark.parametrize("capacity,batch", [(1, 1), (2, 2), (3, 4), (37, 8), (64, 16), (1000, 32),
                                            (2048, 64), (5000, 128), (70000, 256)])
@pytest.mark.parametrize("carrier", ["device"])
def test_random_program_against_oracle(capacity, batch, carrier):
    rng = np.random.default_rng(capacity)
    shapes, types = [(4,), (1,), (1,), (4,)], [np.float32, np.int32, np.float32, np.float32]
    g = M.PriorityReplayBuffer(0.6, capacity, batch, shapes, types, 5, 9)
    o = O.PriorityBufferOracle(0.6, capacity, batch, shapes, types, 5, 9)

    def mk(n):
        return [rng.normal(size=(n, 4)).astype(np.float32), rng.integers(0, 4, (n, 1)).astype(np.int32),
                rng.normal(size=(n, 1)).astype(np.float32), rng.normal(size=(n, 4)).astype(np.float32)]

    # single pushes, then batched pushes (with wrap-around), interleaved with sample/update
    for _ in range(min(capacity, 5)):
        row = [c[0] for c in mk(1)]
        g.insert(*[dev(c) for c in row])
        o.insert(*row)
    same_tree(g, o)
    for n in (max(1, capacity // 3), capacity, max(2, capacity // 2 + 1)):
        b = mk(n)
        g.insert(*[dev(c) for c in b])
        o.insert_batch(b)
        same_tree(g, o)
        for it in range(3):
            u = rng.random(batch).astype(np.float32)
            got = g.sample(0.4, uniforms=dev(u))
            want = o.sample(0.4, uniforms=u)
            same_sample(got, want)
            pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
            if it == 1:  # duplicates inside one update batch: last writer wins
                idx = rng.integers(0, g.info()[0], size=batch)
                idx[: batch // 2] = idx[batch // 2: 2 * (batch // 2)] if batch > 1 else idx[:0]
            else:
                idx = want[0]
            g.update_priorities(dev(idx), dev(pr))
            o.update_priorities(idx, pr)
            same_tree(g, o)
    # the buffer's own generator (no injected uniforms) is the same stream on both sides
    same_sample(g.sample(0.7), o.sample(0.7))
    same_sample(g.sample(0.7), o.sample(0.7))


@pytest.mark.parametrize("batch,types", [(32, [np.float32, np.int32]), (13, [np.float32, np.int32]),
                                         (1, [np.float32, np.int32]), (37, [np.float32, np.uint8]),
                                         (32, [np.float64, np.int16])])
def test_host_carrier_matches_oracle(batch, types):
    """numpy in / numpy out (ragged last CTA: 13, 37; columns that are not whole words: uint8[1], int16[1])"""
    rng = np.random.default_rng(3)
    shapes = [(4,), (1,)]
    g = M.PriorityReplayBuffer(0.6, 500, batch, shapes, types, 1, 2, carrier="numpy")
    o = O.PriorityBufferOracle(0.6, 500, batch, shapes, types, 1, 2)
    for i in range(20):
        row = [rng.normal(size=4).astype(types[0]), np.array([i], types[1])]
        g.insert(*row)
        o.insert(*row)
    b = [rng.normal(size=(700, 4)).astype(types[0]), (np.arange(700) % 100).astype(types[1])[:, None]]
    g.insert(*b)
    o.insert_batch(b)
    for _ in range(4):
        u = rng.random(batch).astype(np.float32)
        got, want = g.sample(0.5, uniforms=u), o.sample(0.5, uniforms=u)
        assert isinstance(got[0], np.ndarray)
        same_sample(got, want)
        pr = (rng.random(batch) * 3).astype(np.float32)
        g.update_priorities(want[0], pr)
        o.update_priorities(want[0], pr)
    same_tree(g, o)


def test_zero_and_negative_priorities_are_clamped():
    g = M.PriorityReplayBuffer(0.6, 16, 4, [(1,)], [np.int32])
    o = O.PriorityBufferOracle(0.6, 16, 4, [(1,)], [np.int32])
    b = [np.arange(16, dtype=np.int32)[:, None]]
    g.insert(*[dev(c) for c in b])
    o.insert_batch(b)
    idx = np.array([0, 1, 2, 3]); pr = np.array([0.0, 5.0, 1e-30, 2.0], np.float32)
    g.update_priorities(dev(idx), dev(pr))
    o.update_priorities(idx, pr)
    same_tree(g, o)
    assert g.tree()[0][16] == np.float32(1e-7)


def test_c2_shape_loop_bit_exact():
    """BASELINE config 2: capacity 1M (2^20 leaves), 40-byte rows, batch 256, alpha .6 beta .4"""
    rng = np.random.default_rng(2)
    cap, batch = 1_000_000, 256
    shapes, types = [(4,), (1,), (1,), (4,)], [np.float32, np.int32, np.float32, np.float32]
    g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 0, 1)
    o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 0, 1)
    obs = rng.normal(size=(cap, 4)).astype(np.float32)
    b = [obs, rng.integers(0, 2, (cap, 1)).astype(np.int32), rng.normal(size=(cap, 1)).astype(np.float32), obs[::-1].copy()]
    g.insert(*[dev(c) for c in b])
    o.insert_batch(b)
    pr0 = (np.abs(rng.normal(size=cap)) + 1e-3).astype(np.float32)
    for s in range(0, cap, 4096):  # initial priorities through the update path itself
        idx = np.arange(s, min(cap, s + 4096))
        g.update_priorities(dev(idx), dev(pr0[idx]))
        o.update_priorities(idx, pr0[idx])
    same_tree(g, o)
    for step in range(60):
        u = rng.random(batch).astype(np.float32)
        got, want = g.sample(0.4, uniforms=dev(u)), o.sample(0.4, uniforms=u)
        same_sample(got, want)
        pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
        g.update_priorities(got[0], dev(pr))
        o.update_priorities(want[0], pr)
    same_tree(g, o)
    gs, _ = g.tree()
    assert np.isfinite(gs[1]) and gs[1] > 0


def test_wide_rows_embed_their_slot():
    """C3-shaped rows (84x84x4 uint8 twice + small columns) at a capacity that fits quickly: the
    gather goes through the TMA ring; every sampled row must carry its own index."""
    cap, batch = 4096, 512
    rb = np.asarray([28224, 4, 4, 1, 28224], dtype=np.int64)
    h = C.c_int64(-1)
    L = _ffi.lib()
    _ffi.check(L.mrl_prb_create(cap, 0.6, 5, rb.ctypes.data, 0, 0, C.byref(h)))
    src = [torch.empty((cap, int(r)), dtype=torch.uint8, device="cuda") for r in rb]
    for c, t in enumerate(src):
        _ffi.check(L.mrl_fill_rows(t.data_ptr(), cap, int(rb[c]), 100 + c, 0))
    _ffi.check(L.mrl_prb_push_batch(h.value, cap, _ffi.ptr_table([t.data_ptr() for t in src]), 0))
    rng = np.random.default_rng(8)
    pr = (np.abs(rng.normal(size=cap)) + 1e-3).astype(np.float32)
    d_all, d_pr = dev(np.arange(cap)), dev(pr)  # keep both alive across the asynchronous call
    _ffi.check(L.mrl_prb_update(h.value, d_all.data_ptr(), d_pr.data_ptr(), cap, 0))
    idx = torch.empty(batch, dtype=torch.int64, device="cuda")
    w = torch.empty(batch, dtype=torch.float32, device="cuda")
    outs = [torch.empty((batch, int(r)), dtype=torch.uint8, device="cuda") for r in rb]
    _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, None, idx.data_ptr(), w.data_ptr(),
                                _ffi.ptr_table([t.data_ptr() for t in outs]), 0))
    torch.cuda.synchronize()
    i = idx.cpu().numpy()
    assert i.min() >= 0 and i.max() < cap and len(np.unique(i)) > batch // 2
    for c in (0, 4):
        got = outs[c].cpu().numpy()
        assert np.array_equal(got[:, :8].copy().view(np.int64)[:, 0], i)
        assert np.array_equal(got, O.fill_rows(cap, int(rb[c]), 100 + c)[i])
    for c in (1, 2, 3):
        assert np.array_equal(outs[c].cpu().numpy(), O.fill_rows(cap, int(rb[c]), 100 + c)[i])
    ww = w.cpu().numpy()
    assert np.all(ww > 0) and np.all(ww <= 1.0)
    _ffi.check(L.mrl_prb_destroy(h.value))


def test_sharded_weights_against_oracle():
    rng = np.random.default_rng(6)
    shapes, types = [(2,)], [np.float32]
    g = M.PriorityReplayBuffer(0.6, 1024, 64, shapes, types, 3, 4)
    o = O.PriorityBufferOracle(0.6, 1024, 64, shapes, types, 3, 4)
    b = [rng.normal(size=(1024, 2)).astype(np.float32)]
    g.insert(*[dev(c) for c in b])
    o.insert_batch(b)
    pr = (rng.random(1024) + 0.01).astype(np.float32)
    g.update_priorities(dev(np.arange(1024)), dev(pr))
    o.update_priorities(np.arange(1024), pr)
    st = host(g.root_stats())
    assert np.array_equal(st, o.root_stats())
    others = np.array([[st[0] * 2.5, st[1] * 0.5, 777.0, 3.0], [st[0] * 0.1, st[1] * 4, 10.0, 1.0]], np.float32)
    gstats = np.concatenate([others[:1], st[None], others[1:]], 0)
    u = rng.random(64).astype(np.float32)
    got = g.sample(0.4, uniforms=dev(u), gstats=dev(gstats))
    want = o.sample(0.4, uniforms=u, gstats=gstats)
    same_sample(got, want)
    # option "shard_weights" = 1: weights against the true sampling probability p / (world * S_rank)
    try:
        _ffi.set_option("shard_weights", 1)
        O.set_shard_weights(1)
        got1 = g.sample(0.4, uniforms=dev(u), gstats=dev(gstats))
        want1 = o.sample(0.4, uniforms=u, gstats=gstats)
        same_sample(got1, want1)
        assert np.array_equal(host(got1[0]), host(got[0])) and not np.array_equal(host(got1[1]), host(got[1]))
    finally:
        _ffi.set_option("shard_weights", 0)
        O.set_shard_weights(0)


def test_checkpoint_round_trip():
    """state_dict()/load_state_dict(): a restored buffer continues exactly like the original"""
    rng = np.random.default_rng(9)
    shapes, types = [(3,), (1,)], [np.float32, np.int32]
    g = M.PriorityReplayBuffer(0.6, 3000, 32, shapes, types, 4, 5)
    b = [rng.normal(size=(2500, 3)).astype(np.float32), np.arange(2500, dtype=np.int32)[:, None]]
    g.insert(*[dev(c) for c in b])
    g.update_priorities(dev(np.arange(2500)), dev((rng.random(2500) * 3 + 0.01).astype(np.float32)))
    g.sample(0.4)  # advances the generator
    sd = g.state_dict()
    h = M.PriorityReplayBuffer(0.6, 3000, 32, shapes, types, 4, 5)
    h.load_state_dict(sd)
    assert h.info() == g.info() and h.max_priority() == g.max_priority()
    for _ in range(3):
        a, c = g.sample(0.4), h.sample(0.4)
        for x, y in zip(a, c):
            assert torch.equal(x, y)
        pr = dev((rng.random(32) + 0.01).astype(np.float32))
        g.update_priorities(a[0], pr)
        h.update_priorities(c[0], pr)
    row = [dev(np.ones(3, np.float32)), dev(np.array([7], np.int32))]
    g.insert(*row)
    h.insert(*row)
    gs, gm = g.tree()
    hs, hm = h.tree()
    assert np.array_equal(gs[1:], hs[1:]) and np.array_equal(gm[1:], hm[1:])


@pytest.mark.parametrize("batch", [1, 37, 256, 300])
def test_host_update_inline_batch_matches_staged_and_oracle(batch):
    """mrl_prb_update_host: batches of <= 256 ride in the kernel's parameter block (no copy); larger ones
    and inline_update=0 stage through device memory.  Same tree either way, duplicates and clamped
    priorities included, bit-exact against the oracle."""
    from mindrl_b200 import _ffi
    rng = np.random.default_rng(batch)
    cap = 70000  # 2^17 leaves: the subtree-partitioned kernel
    shapes, types = [(2,)], [np.float32]
    bufs = [M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 3, 4, carrier="numpy") for _ in range(2)]
    o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 3, 4)
    rows = [rng.normal(size=(cap, 2)).astype(np.float32)]
    for g in bufs:
        g.insert(*rows)
    o.insert_batch(rows)
    try:
        for it in range(6):
            idx = rng.integers(0, cap, size=batch).astype(np.int64)
            if batch > 1 and it % 2:
                idx[: batch // 2] = idx[batch // 2: 2 * (batch // 2)]  # duplicates: last writer wins
            pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
            if it == 3:
                pr[::3] = 0.0
                pr[1::5] = -1.0
            _ffi.set_option("inline_update", 1)
            bufs[0].update_priorities(idx, pr)
            _ffi.set_option("inline_update", 0)
            bufs[1].update_priorities(idx, pr)
            o.update_priorities(idx, pr)
            for g in bufs:
                same_tree(g, o)
                assert g.max_priority() == o.max_priority()
    finally:
        _ffi.set_option("inline_update", 1)


@pytest.mark.parametrize("pdl", [1, 0])
def test_sample_update_captured_in_a_cuda_graph(pdl):
    """The device entry points are pure stream work (cursors are host integers, no hidden syncs), so a
    learner can capture sample -> update_priorities once and replay it; each replay is checked bit for bit
    against the oracle.  (Not capturable by design: draws from the internal generator and insert(), whose
    counters / cursors advance on the host.)"""
    rng = np.random.default_rng(11)
    cap, batch = 5000, 64
    shapes, types = [(4,), (1,)], [np.float32, np.int32]
    _ffi.set_option("pdl", pdl)
    try:
        g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 0, 1)
        o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 0, 1)
        b = [rng.normal(size=(cap, 4)).astype(np.float32), rng.integers(0, 9, (cap, 1)).astype(np.int32)]
        g.insert(*[dev(c) for c in b])
        o.insert_batch(b)
        pr0 = (np.abs(rng.normal(size=cap)) + 1e-3).astype(np.float32)
        g.update_priorities(dev(np.arange(cap)), dev(pr0))
        o.update_priorities(np.arange(cap), pr0)
        u_dev = torch.zeros(batch, dtype=torch.float32, device="cuda")
        p_dev = torch.ones(batch, dtype=torch.float32, device="cuda")
        # one eager step first (one-off function attributes are set outside the capture), mirrored on the oracle
        u = rng.random(batch).astype(np.float32)
        pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
        u_dev.copy_(torch.from_numpy(u)), p_dev.copy_(torch.from_numpy(pr))
        got = g.sample(0.4, uniforms=u_dev)
        g.update_priorities(got[0], p_dev)
        want = o.sample(0.4, uniforms=u)
        o.update_priorities(want[0], pr)
        same_sample(got, want)
        torch.cuda.synchronize()
        launches0 = _ffi.launch_count()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            got = g.sample(0.4, uniforms=u_dev)
            g.update_priorities(got[0], p_dev)
        assert _ffi.launch_count() - launches0 == 2  # captured, not run
        same_tree(g, o)  # capture did not execute anything
        for _ in range(5):
            u = rng.random(batch).astype(np.float32)
            pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
            u_dev.copy_(torch.from_numpy(u)), p_dev.copy_(torch.from_numpy(pr))
            graph.replay()
            torch.cuda.synchronize()
            want = o.sample(0.4, uniforms=u)
            o.update_priorities(want[0], pr)
            same_sample(got, want)
            same_tree(g, o)
    finally:
        _ffi.set_option("pdl", 1)


@pytest.mark.parametrize("option,value", [("tree_prefetch", 0), ("host_zero_copy", 0), ("pdl", 0), ("inline_update", 0)])
def test_options_only_select_among_equal_paths(option, value):
    """every tunable of the PER path changes how the work is done, never the result: the same program with
    the option flipped is bit-identical to the oracle (device and host carriers)"""
    rng = np.random.default_rng(5)
    cap, batch = 40000, 96
    shapes, types = [(4,), (1,)], [np.float32, np.int32]
    _ffi.set_option(option, value)
    try:
        for carrier in ("device", "numpy"):
            g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 3, 4, carrier=carrier)
            o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 3, 4)
            b = [rng.normal(size=(cap, 4)).astype(np.float32), rng.integers(0, 9, (cap, 1)).astype(np.int32)]
            g.insert(*([dev(c) for c in b] if carrier == "device" else b))
            o.insert_batch(b)
            for _ in range(4):
                u = rng.random(batch).astype(np.float32)
                got = g.sample(0.4, uniforms=dev(u) if carrier == "device" else u)
                want = o.sample(0.4, uniforms=u)
                same_sample(got, want)
                pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
                g.update_priorities(got[0], dev(pr) if carrier == "device" else pr)
                o.update_priorities(want[0], pr)
            same_tree(g, o)
    finally:
        _ffi.set_option(option, 1)


def test_full_size_c3_indices_weights_rows():
    """BASELINE config 3 at full size: capacity 1M, two 84x84x4 uint8 columns (2 x 28.2 GB: row offsets
    need 64 bits) + action / reward / done, batch 512.  Indices and weights bit-exact against an oracle
    buffer with the same tree (its rows are 1 byte: the tree does not depend on them); every gathered wide
    row must be the row the fill function defines for that index, whole, and carry its index in bytes 0..7."""
    free, _ = torch.cuda.mem_get_info()
    if free < 70 << 30:
        pytest.skip("needs 70 GB of free device memory")
    cap, batch = 1_000_000, 512
    rb = np.asarray([28224, 4, 4, 1, 28224], dtype=np.int64)
    h = C.c_int64(-1)
    L = _ffi.lib()
    _ffi.check(L.mrl_prb_create(cap, 0.6, 5, rb.ctypes.data, 3, 4, C.byref(h)))
    try:
        cols = (C.c_void_p * 5)()
        _ffi.check(L.mrl_prb_columns(h.value, cols))
        for c in range(5):
            _ffi.check(L.mrl_fill_rows(cols[c], cap, int(rb[c]), 100 + c, 0))
        _ffi.check(L.mrl_prb_push_batch(h.value, cap, cols, 0))  # rows onto themselves: ring full, leaves = max priority
        o = O.PriorityBufferOracle(0.6, cap, batch, [(1,)], [np.uint8], 3, 4)
        o.insert_batch([np.zeros((cap, 1), np.uint8)])
        rng = np.random.default_rng(33)
        pr = (np.abs(rng.normal(size=cap)) + 1e-3).astype(np.float32)
        d_idx, d_pr = dev(np.arange(cap)), dev(pr)
        for s in range(0, cap, 65536):
            e = min(cap, s + 65536)
            _ffi.check(L.mrl_prb_update(h.value, d_idx[s:e].data_ptr(), d_pr[s:e].data_ptr(), e - s, 0))
            o.update_priorities(np.arange(s, e), pr[s:e])
        idx = torch.empty(batch, dtype=torch.int64, device="cuda")
        w = torch.empty(batch, dtype=torch.float32, device="cuda")
        outs = [torch.empty((batch, int(r)), dtype=torch.uint8, device="cuda") for r in rb]
        tab = _ffi.ptr_table([t.data_ptr() for t in outs])
        seen_high = False
        for step in range(4):
            u = rng.random(batch).astype(np.float32)
            du = dev(u)
            _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, du.data_ptr(), idx.data_ptr(), w.data_ptr(), tab, 0))
            torch.cuda.synchronize()
            want = o.sample(0.4, uniforms=u)
            i = idx.cpu().numpy()
            assert np.array_equal(i, want[0]), "indices differ"
            assert np.array_equal(w.cpu().numpy().view(np.uint32), want[1].view(np.uint32)), "weights differ"
            seen_high = seen_high or bool((i * 28224 >= 1 << 32).any())
            for c in range(5):
                got = outs[c].cpu().numpy()
                assert np.array_equal(got, O.fill_rows_at(i, int(rb[c]), 100 + c)), f"column {c} rows differ"
            assert np.array_equal(outs[0].cpu().numpy()[:, :8].copy().view(np.int64)[:, 0], i)
            p2 = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
            dp2 = dev(p2)
            _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), dp2.data_ptr(), batch, 0))
            o.update_priorities(want[0], p2)
            torch.cuda.synchronize()
        assert seen_high  # rows beyond the 4 GiB offset were gathered
    finally:
        _ffi.check(L.mrl_prb_destroy(h.value))


@pytest.mark.parametrize("pattern", ["one_subtree", "two_subtrees_one_cta", "mixed"])
def test_update_shares_that_overflow_a_warp(pattern):
    """the update kernel gives each level-11 subtree to a warp (<= 32 entries); a crowded subtree must send the
    CTA down the general paths -- every time (the decision once raced with the counting)"""
    rng = np.random.default_rng(21)
    cap, batch = 1 << 20, 256
    shapes, types = [(1,)], [np.int32]
    g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 0, 1)
    o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 0, 1)
    b = [np.arange(cap, dtype=np.int32)[:, None]]
    g.insert(*[dev(c) for c in b])
    o.insert_batch(b)
    for rep in range(20):
        if pattern == "one_subtree":          # 512 leaves per subtree: 200 entries in one of them
            idx = rng.integers(0, 512, size=200) + 512 * int(rng.integers(0, 2048))
        elif pattern == "two_subtrees_one_cta":  # subtrees t and t + 128 belong to the same CTA, different warps
            t = int(rng.integers(0, 128))
            idx = np.concatenate([rng.integers(0, 512, size=40) + 512 * t, rng.integers(0, 512, size=20) + 512 * (t + 128)])
        else:
            idx = np.concatenate([rng.integers(0, cap, size=150), rng.integers(0, 512, size=33) + 512 * 7])
        if rep % 2:  # batches above 1024 entries use a warp per subtree: crowd one while the rest is sparse
            idx = np.concatenate([idx, rng.integers(0, cap, size=1500)])
        idx = idx[idx < cap].astype(np.int64)
        pr = (np.abs(rng.normal(size=idx.size)) + 1e-3).astype(np.float32)
        g.update_priorities(dev(idx), dev(pr))
        o.update_priorities(idx, pr)
        same_tree(g, o)


def test_c5_per_rank_shape_bit_exact():
    """BASELINE config 5's per-rank shape on ONE GPU: capacity 1M (2^20 leaves), 40-byte rows, sample(4096) +
    update_priorities(4096) -- the batch the sharded bench runs per rank.  Indices, weights, rows after every step
    and the whole tree at the end, bit for bit against the oracle; both the internal generator and injected draws."""
    rng = np.random.default_rng(55)
    cap, batch = 1_000_000, 4096
    shapes, types = [(4,), (1,), (1,), (4,)], [np.float32, np.int32, np.float32, np.float32]
    g = M.PriorityReplayBuffer(ALPHA := 0.6, cap, batch, shapes, types, 7, 0)
    o = O.PriorityBufferOracle(ALPHA, cap, batch, shapes, types, 7, 0)
    rows = [rng.normal(size=(cap, 4)).astype(np.float32), rng.integers(0, 4, (cap, 1)).astype(np.int32),
            rng.normal(size=(cap, 1)).astype(np.float32), rng.normal(size=(cap, 4)).astype(np.float32)]
    g.insert(*[dev(c) for c in rows])
    o.insert_batch(rows)
    pr0 = (np.abs(rng.normal(size=cap)) + 1e-3).astype(np.float32)
    g.update_priorities(dev(np.arange(cap)), dev(pr0))
    o.update_priorities(np.arange(cap), pr0)
    for step in range(6):
        u = None if step % 2 else rng.random(batch).astype(np.float32)
        got = g.sample(0.4, uniforms=None if u is None else dev(u))
        want = o.sample(0.4, uniforms=u)
        same_sample(got, want)
        pr = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
        if step == 3:  # duplicates inside one 4096-batch: the last writer wins
            idx = host(got[0]).copy()
            idx[1000:2000] = idx[:1000]
            g.update_priorities(dev(idx), dev(pr))
            o.update_priorities(idx, pr)
        else:
            g.update_priorities(got[0], dev(pr))
            o.update_priorities(want[0], pr)
    same_tree(g, o)


def test_sample_of_an_empty_buffer_is_an_error():
    """an empty buffer used to return index 0 and NaN weights silently"""
    g = M.PriorityReplayBuffer(0.6, 100, 8, [(2,)], [np.float32])
    with pytest.raises(_ffi.MindRLError) as e:
        g.sample(0.4)
    assert e.value.code == _ffi.MRL_ERR_STATE
    o = O.PriorityBufferOracle(0.6, 100, 8, [(2,)], [np.float32])
    with pytest.raises(RuntimeError):
        o.sample(0.4)


def test_indices_stay_below_size_and_match_their_rows():
    """a mass that rounds past the total used to surface the raw (padding / unfilled) leaf as the index
    while the row came from a clamped slot.  Now index, weight and row all belong to the newest valid row.  Draws
    of exactly 1.0 - 2^-24 in the last stratum push the mass to (and, by rounding, past) the total."""
    rng = np.random.default_rng(8)
    for cap, fill in ((1000, 700), (1024, 1024), (5, 3)):
        batch = 16
        g = M.PriorityReplayBuffer(0.7, cap, batch, [(1,)], [np.int64])
        o = O.PriorityBufferOracle(0.7, cap, batch, [(1,)], [np.int64])
        rows = [np.arange(fill, dtype=np.int64)[:, None]]
        g.insert(*[dev(c) for c in rows])
        o.insert_batch(rows)
        pr = (rng.random(fill) * 3 + 1e-3).astype(np.float32)
        g.update_priorities(dev(np.arange(fill)), dev(pr))
        o.update_priorities(np.arange(fill), pr)
        for u_last in (np.float32(1.0) - np.float32(2.0 ** -24), np.float32(1.0)):  # 1.0 is outside [0,1): worst case
            u = rng.random(batch).astype(np.float32)
            u[-1] = u_last
            got = g.sample(0.5, uniforms=dev(u))
            want = o.sample(0.5, uniforms=u)
            same_sample(got, want)
            idx = host(got[0])
            assert np.all(idx < fill) and np.all(idx >= 0)
            assert np.array_equal(host(got[2]).reshape(-1), idx)
            assert np.all(np.isfinite(host(got[1])))


def test_host_update_rejects_indices_outside_the_buffer():
    """mrl_prb_update_host wrote out of bounds for an index >= capacity"""
    g = M.PriorityReplayBuffer(0.6, 5000, 8, [(2,)], [np.float32], carrier="numpy")
    g.insert(np.zeros((5000, 2), np.float32))
    before = g.tree()
    for bad in (5000, 1 << 20, -1, 1 << 40):
        idx = np.array([1, 2, bad, 3], dtype=np.int64)
        with pytest.raises(ValueError):
            g.update_priorities(idx, np.ones(4, np.float32))
    after = g.tree()
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[1], after[1])
    big = np.arange(600, dtype=np.int64)  # staged (not inline) path
    big[-1] = 5000
    with pytest.raises(ValueError):
        g.update_priorities(big, np.ones(600, np.float32))


@pytest.mark.parametrize("batch", [1, 2, 37, 512, 592, 593, 2048])
@pytest.mark.parametrize("row", [28224, 4096, 2064])
def test_fused_sample_gather_equals_the_two_kernel_path_and_the_oracle(batch, row):
    """Atari-shaped rows: prb_sample_gather_kernel (descent + DMA ring in one launch, batches that fit it) against
    prb_sample_kernel + gather_bulk_kernel (option fused_gather = 0) and the oracle -- indices, weight bits, every
    byte of every column; a partly filled ring; injected draws and the internal generator."""
    rng = np.random.default_rng(batch * 31 + row)
    cap, fill = 6000, 5000
    rb = [row, 4, 4, 1, row]
    shapes, types = [(r,) for r in rb], [np.uint8] * 5
    rows = [O.fill_rows(fill, r, 50 + c) for c, r in enumerate(rb)]
    pr = (np.abs(rng.normal(size=fill)) + 1e-3).astype(np.float32)
    o = O.PriorityBufferOracle(0.6, cap, batch, shapes, types, 5, 6)
    o.insert_batch(rows)
    o.update_priorities(np.arange(fill), pr)
    bufs = {}
    try:
        for fused in (1, 0):
            _ffi.set_option("fused_gather", fused)
            g = M.PriorityReplayBuffer(0.6, cap, batch, shapes, types, 5, 6)
            g.insert(*[dev(c) for c in rows])
            g.update_priorities(dev(np.arange(fill)), dev(pr))
            bufs[fused] = g
        for step in range(3):
            u = None if step == 1 else rng.random(batch).astype(np.float32)
            want = o.sample(0.4, uniforms=u)
            newp = (np.abs(rng.normal(size=batch)) + 1e-3).astype(np.float32)
            for fused, g in bufs.items():
                _ffi.set_option("fused_gather", fused)
                before = _ffi.launch_count()
                got = g.sample(0.4, uniforms=None if u is None else dev(u))
                launches = _ffi.launch_count() - before
                same_sample(got, want)
                assert np.array_equal(host(got[2])[:, :8].copy().view(np.int64)[:, 0], want[0])  # row carries its slot
                if fused and batch <= 592:
                    assert launches == 1, launches  # one kernel: descent and copy
                elif not fused:
                    assert launches >= 2
                g.update_priorities(got[0], dev(newp))
            o.update_priorities(want[0], newp)
        for g in bufs.values():
            same_tree(g, o)
    finally:
        _ffi.set_option("fused_gather", 1)
