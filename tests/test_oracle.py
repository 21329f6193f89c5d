"""The CPU oracle against everything that can pin it: the reference's known-answer vectors, golden
vectors produced by running the reference's own Python (tests/golden/make_golden.py), a
numpy twin of the reference loop, and the behavioural expectations of the reference's (skipped)
buffer tests."""
import os

import numpy as np
import pytest

from oracle import oracle as O


# ---- scans -------------------------------------------------------------------------------
def test_reference_known_answers():
    # tests/st/test_discounted_return.py:39-45
    r = np.array([[1], [1], [1], [1]], np.float32)
    d = np.array([[False], [False], [True], [False]])
    out = O.discounted_return(r, d, np.array([2.0], np.float32), 0.99)
    assert np.allclose(out, np.array([[2.9701], [1.99], [1.0], [2.98]], np.float32))
    # :47-53
    r = np.array([1, 1, 1, 1], np.float32)
    d = np.array([False, True, False, True])
    out = O.discounted_return(r, d, np.float32(2.0), 0.99)
    assert np.allclose(out, np.array([1.99, 1, 1.99, 1.0], np.float32))


def test_discounted_return_golden_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "discounted_return_ref.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) >= 20
    for n in names:
        out = O.discounted_return(g[n + "/reward"], g[n + "/done"], g[n + "/last"], float(g[n + "/gamma"]))
        assert out.shape == g[n + "/out"].shape
        assert np.array_equal(out.view(np.uint32), g[n + "/out"].view(np.uint32)), n
        twin = O.discounted_return_numpy_loop(g[n + "/reward"], g[n + "/done"], g[n + "/last"],
                                              float(g[n + "/gamma"]))
        assert np.array_equal(twin.view(np.uint32), g[n + "/out"].view(np.uint32)), n


def test_ppo_golden_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "ppo_scans_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        gamma, lam = float(g[n + "/gamma"]), float(g[n + "/lam"])
        dr = O.discounted_return(g[n + "/reward"], None, None, gamma, layout=1)
        assert np.array_equal(dr.view(np.uint32), g[n + "/discounted_r"].view(np.uint32)), n
        adv, ret = O.gae(g[n + "/reward"], g[n + "/value"], g[n + "/next_value"], None, None,
                         gamma, lam, layout=1)
        assert np.array_equal(adv.view(np.uint32), g[n + "/advantage"].view(np.uint32)), n
        assert np.array_equal(ret, adv + g[n + "/value"])


def test_layouts_agree():
    rng = np.random.default_rng(0)
    T, B = 37, 5
    r = rng.normal(size=(T, B)).astype(np.float32)
    v = rng.normal(size=(T, B)).astype(np.float32)
    lv = rng.normal(size=(B,)).astype(np.float32)
    d = rng.random((T, B)) < 0.2
    a0 = O.discounted_return(r, d, lv, 0.97, layout=0)
    a1 = O.discounted_return(r.T.copy(), d.T.copy(), lv, 0.97, layout=1)
    assert np.array_equal(a0, a1.T)
    g0 = O.gae(r, v, None, lv, d, 0.99, 0.95, layout=0)
    g1 = O.gae(r.T.copy(), v.T.copy(), None, lv, d.T.copy(), 0.99, 0.95, layout=1)
    assert np.array_equal(g0[0], g1[0].T) and np.array_equal(g0[1], g1[1].T)


def test_mappo_form_matches_python_restatement():
    # mappo.py:478-497 written out with numpy float32 scalars, time-major with masks
    rng = np.random.default_rng(1)
    T, B = 25, 4
    r = rng.normal(size=(T, B)).astype(np.float32)
    val = rng.normal(size=(T + 1, B)).astype(np.float32)
    mask = (rng.random((T, B)) > 0.15).astype(np.float32)
    gamma, lam = np.float32(0.99), np.float32(0.95)
    gae = np.zeros(B, np.float32)
    ret = np.zeros((T, B), np.float32)
    for step in range(T - 1, -1, -1):
        delta = r[step] + gamma * val[step + 1] * mask[step] - val[step]
        gae = delta + gamma * lam * mask[step] * gae
        ret[step] = gae + val[step]
    adv_o, ret_o = O.gae(r, val[:T], None, val[T], 1.0 - mask, float(gamma), float(lam), layout=0)
    assert np.array_equal(ret_o, ret)
    assert np.array_equal(adv_o, ret - val[:T]) or np.allclose(adv_o, ret - val[:T], rtol=1e-6, atol=1e-6)


def test_affine_scan_is_vtrace_recurrence():
    rng = np.random.default_rng(2)
    T, B = 20, 3
    delta = rng.normal(size=(T, B)).astype(np.float32)
    disc = np.full((T, B), 0.99, np.float32)
    cs = np.minimum(1.0, rng.random((T, B)) * 2).astype(np.float32)
    masks = (rng.random((T, B)) > 0.1).astype(np.float32)
    acc = np.zeros(B, np.float32)
    ref = np.zeros((T, B), np.float32)
    for i in range(T - 1, -1, -1):  # vtrace.py:88-93
        acc = delta[i] + disc[i] * cs[i] * acc * masks[i]
        ref[i] = acc
    out = O.affine_scan(delta, disc * cs * masks, None, layout=0)
    assert np.allclose(out, ref, rtol=1e-6, atol=1e-6)


def test_vtrace_golden_through_oracle_scan(golden_dir):
    """vtrace.py:84-99 run by the reference's own code: the recurrence is the oracle's affine scan"""
    g = np.load(os.path.join(golden_dir, "vtrace_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        lr, disc, r, v, boot, m = (g[f"{n}/{k}"] for k in ("log_rhos", "discounts", "rewards", "values", "boot", "masks"))
        rhos = np.minimum(np.exp(lr), np.float32(1.0))
        cs = np.minimum(rhos, np.float32(1.0))
        v1 = np.concatenate([v[1:], boot[None]], 0)
        deltas = rhos * (r + disc * v1 - v)
        acc = O.affine_scan(deltas, disc * cs * m, None, layout=0)
        assert np.array_equal((acc + v).view(np.uint32), g[f"{n}/vs"].view(np.uint32)), n


# ---- arithmetic --------------------------------------------------------------------------
def test_detpow_equals_correctly_rounded_pow():
    rng = np.random.default_rng(3)
    x = (np.abs(rng.normal(size=2_000_000)) + 1e-3).astype(np.float32)
    for y in (0.6, -0.4, 0.5, 1.0, -1.0, 0.37):
        a, b = O.detpow(x, np.float32(y)), O.libm_pow(x, np.float32(y))
        ulp = np.abs(a.view(np.int32).astype(np.int64) - b.view(np.int32))
        assert ulp.max() <= 1
        assert (ulp != 0).mean() < 1e-5
    x = (10 ** rng.uniform(-30, 30, size=1_000_000)).astype(np.float32)
    y = rng.uniform(-3, 3, size=x.size).astype(np.float32)
    a, b = O.detpow(x, y), O.libm_pow(x, y)
    fin = np.isfinite(b) & (b > 0)
    ulp = np.abs(a[fin].view(np.int32).astype(np.int64) - b[fin].view(np.int32))
    assert ulp.max() <= 1 and (ulp != 0).mean() < 1e-5
    assert np.array_equal(a[~fin], b[~fin])


def test_detpow_edges():
    f = O.lib().orc_detpowf
    assert f(0.0, 0.6) == 0.0 and f(0.0, -0.4) == np.inf and f(5.0, 0.0) == 1.0
    assert f(1.0, 123.0) == 1.0 and f(np.inf, -1.0) == 0.0 and np.isnan(f(-1.0, 0.5))
    assert f(1e-45, 0.5) > 0  # subnormal input


def test_uniform01_range():
    u = np.array([O.uniform01(42, i) for i in range(20000)], np.float32)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02


# ---- uniform buffer ----------------------------------------------------------------------
def _model_append(state, cap, n):
    """pure-python cursor model, one row at a time"""
    count, head, order = state
    for _ in range(n):
        if head == 0 and count < cap:
            slot = count
            count += 1
        else:
            slot = head
            head = (head + 1) % cap
        order.append(slot)
    return count, head, order


def test_uniform_cursor_and_fifo():
    cap = 10
    b = O.UniformBufferOracle(4, cap, [(2,), (1,)], [np.float32, np.int32])
    state = (0, 0, [])
    written = {}
    step = 0
    for n in (1, 3, 4, 5, 2, 10, 7, 1):
        a = (np.arange(step, step + n, dtype=np.float32)[:, None] * np.ones((1, 2), np.float32))
        c = np.arange(step, step + n, dtype=np.int32)[:, None]
        b.insert([a, c])
        state = _model_append(state, cap, n)
        for j, slot in enumerate(state[2][-n:]):
            written[slot] = step + j
        step += n
        count, head = b.state()
        assert (count, head) == state[:2]
        cols = b.columns()
        for slot, val in written.items():
            assert cols[0][slot, 0] == val and cols[1][slot, 0] == val
        # logical order: 0 = oldest ... count-1 = newest
        vals = [int(b.get_item(i)[1][0]) for i in range(count)]
        assert vals == sorted(vals) and vals[-1] == step - 1
        assert int(b.get_item(-1)[1][0]) == step - 1
    with pytest.raises(IndexError):
        b.get_item(cap)


def test_uniform_overlong_batch_and_reset():
    cap = 8
    b = O.UniformBufferOracle(4, cap, [(1,)], [np.int64])
    b.insert([np.arange(3, dtype=np.int64)])
    b.insert([np.arange(100, 120, dtype=np.int64)])  # 20 rows into capacity 8
    count, head = b.state()
    assert count == cap
    vals = [int(b.get_item(i)[0][0]) for i in range(count)]
    assert vals == list(range(112, 120))
    b.reset()
    assert b.state() == (0, head)  # head untouched (uniform_replay_buffer.py:141-150)


def test_uniform_sample_rows_match_slots():
    cap = 50
    b = O.UniformBufferOracle(16, cap, [(3,), (1,)], [np.float32, np.int32], seed=5)
    with pytest.raises(RuntimeError):
        b.sample()
    b.insert([np.repeat(np.arange(30, dtype=np.float32)[:, None], 3, 1), np.arange(30, dtype=np.int32)[:, None]])
    slots, (s0, s1) = b.sample()
    assert slots.min() >= 0 and slots.max() < 30
    assert np.array_equal(s1[:, 0], slots) and np.array_equal(s0[:, 0], slots.astype(np.float32))
    slots2, _ = b.sample()
    assert not np.array_equal(slots, slots2)  # the reference test's only assertion (:85)
    g = b.gather(slots, threads=4)
    assert np.array_equal(g[1][:, 0], slots)


# ---- priority buffer ---------------------------------------------------------------------
def _check_tree(b):
    s, m = b.tree()
    cap2 = b.info()[2]
    for i in range(1, cap2):
        assert s[i] == np.float32(s[2 * i] + s[2 * i + 1])
        assert m[i] == min(m[2 * i], m[2 * i + 1])


def test_priority_reference_behaviour():
    # tests/st/test_priority_replay_buffer.py:36-66 (skipped upstream), same sizes and checks
    capacity, batch = 200, 32
    b = O.PriorityBufferOracle(1.0, capacity, batch, [(17,), (6,)], [np.float32, np.int32], 0, 42)
    for i in range(100):
        b.insert(np.ones(17, np.float32) * i, np.ones(6, np.int32) * i)
    idx, w, states, actions = b.sample(1.0)
    assert np.all(idx < 100)
    assert np.allclose(states, np.broadcast_to(idx.reshape(-1, 1), states.shape))
    assert np.allclose(actions, np.broadcast_to(idx.reshape(-1, 1), actions.shape))
    assert np.allclose(w, 1.0)  # all priorities equal -> all weights 1
    b.update_priorities(idx, np.ones(batch, np.float32) * 1e-7)
    idx2, _, states2, _ = b.sample(1.0)
    assert np.all(idx2 < 100) and np.all(idx != idx2)
    assert np.allclose(states2, np.broadcast_to(idx2.reshape(-1, 1), states2.shape))
    _check_tree(b)


def test_priority_tree_matches_naive_model():
    rng = np.random.default_rng(4)
    cap = 37  # not a power of two
    b = O.PriorityBufferOracle(0.6, cap, 8, [(1,)], [np.int64], 1, 2)
    leaves = np.zeros(64, np.float32)
    maxp = np.float32(1.0)
    head = -1
    for step in range(120):
        head = (head + 1) % cap
        b.insert(np.array([step], np.int64))
        leaves[head] = maxp
        if step % 7 == 3:
            idx = rng.integers(0, min(step + 1, cap), size=5)
            pr = (np.abs(rng.normal(size=5)) + 1e-3).astype(np.float32)
            b.update_priorities(idx, pr)
            for i, p in zip(idx, pr):
                pa = O.detpow(np.float32(p), np.float32(0.6)).reshape(())
                leaves[i] = pa
                maxp = max(maxp, np.float32(pa))
    s, m = b.tree()
    assert np.array_equal(s[64:128], leaves)
    assert b.max_priority() == maxp
    _check_tree(b)
    size, hd, cap2 = b.info()
    assert (size, hd, cap2) == (cap, head, 64)


def test_priority_duplicates_last_writer_wins_and_clamp():
    b = O.PriorityBufferOracle(1.0, 16, 4, [(1,)], [np.int32])
    for i in range(16):
        b.insert(np.array([i], np.int32))
    b.update_priorities(np.array([3, 5, 3, 3]), np.array([9.0, 2.0, 4.0, 0.5], np.float32))
    s, _ = b.tree()
    assert s[16 + 3] == np.float32(0.5) and s[16 + 5] == np.float32(2.0)
    assert b.max_priority() == np.float32(9.0)  # max over all processed priorities
    b.update_priorities(np.array([1]), np.array([0.0], np.float32))
    s, _ = b.tree()
    assert s[16 + 1] == np.float32(1e-7)  # non-positive p**alpha is clamped


def test_priority_sampling_is_proportional_and_stratified():
    rng = np.random.default_rng(5)
    n = 64
    b = O.PriorityBufferOracle(1.0, n, 16, [(1,)], [np.int32])
    for i in range(n):
        b.insert(np.array([i], np.int32))
    pr = (rng.random(n) + 0.05).astype(np.float32)
    b.update_priorities(np.arange(n), pr)
    s, _ = b.tree()
    total = s[1]
    cum = np.cumsum(s[n:2 * n].astype(np.float64))
    hits = np.zeros(n)
    for _ in range(400):
        u = rng.random(16).astype(np.float32)
        idx, w, rows = b.sample(0.4, uniforms=u)
        mass = (u.astype(np.float64) + np.arange(16)) * total / 16
        expect = np.searchsorted(cum, mass, side="left")
        assert np.mean(idx == np.minimum(expect, n - 1)) > 0.95  # up to fp32 rounding at edges
        assert np.array_equal(rows[:, 0], idx)
        np.add.at(hits, idx, 1)
        assert w.max() <= 1.0 + 1e-6 and w.min() > 0
    freq = hits / hits.sum()
    assert np.abs(freq - pr / pr.sum()).max() < 0.01


def test_priority_sharded_weights():
    b = O.PriorityBufferOracle(1.0, 32, 8, [(1,)], [np.int32])
    for i in range(32):
        b.insert(np.array([i], np.int32))
    b.update_priorities(np.arange(32), np.linspace(0.1, 2.0, 32).astype(np.float32))
    u = np.linspace(0.05, 0.95, 8).astype(np.float32)
    st = b.root_stats()
    i1, w1 = b.sample(0.5, uniforms=u, gather=False)
    i2, w2 = b.sample(0.5, uniforms=u, gather=False, gstats=st[None])
    assert np.array_equal(i1, i2) and np.array_equal(w1, w2)  # world of one == local
    other = np.array([st[0] * 3, st[1] / 2, 64.0, 1.0], np.float32)
    i3, w3 = b.sample(0.5, uniforms=u, gather=False, gstats=np.stack([st, other]))
    assert np.array_equal(i1, i3) and not np.array_equal(w1, w3) and w3.max() <= 1.0


def test_reservoir_oracle_keeps_an_unbiased_sample():
    """Vitter's property the reference's docstring promises ("keeps an 'unbiased' sample of previous
    iterations", reservoir_replay_buffer.py:27-31): after N offers every transition is in the reservoir with
    probability capacity / N, whatever its age."""
    cap, n_offers, trials = 20, 200, 600
    hits = np.zeros(n_offers)
    for trial in range(trials):
        o = O.ReservoirBufferOracle(cap, 8, [(1,)], [np.int64], seed0=trial, seed1=7)
        for i in range(n_offers):
            o.push(np.asarray([i]))
        hits[o.cols[0][:, 0]] += 1
        assert len(set(o.cols[0][:, 0].tolist())) == cap  # distinct transitions
    p = hits / trials
    assert abs(p.mean() - cap / n_offers) < 1e-12
    # thirds of the stream by age: each within 4 sigma of capacity / N
    for part in np.array_split(p, 4):
        sigma = np.sqrt((cap / n_offers) * (1 - cap / n_offers) / (trials * len(part)))
        assert abs(part.mean() - cap / n_offers) < 4 * sigma + 1e-3
    slots, cols = o.sample()
    assert len(slots) == 8 and len(set(slots.tolist())) == 8
    assert np.array_equal(cols[0], o.cols[0][slots])


def test_fill_rows_at_is_a_row_subset_of_fill_rows():
    """the full-size buffer tests verify gathered rows one by one: the per-row filler must agree with the bulk one"""
    for rb in (1, 3, 8, 37, 28224):
        full = O.fill_rows(300, rb, 100 + rb)
        rows = np.array([0, 299, 17, 17, 123], dtype=np.int64)
        assert np.array_equal(O.fill_rows_at(rows, rb, 100 + rb), full[rows])
        if rb >= 8:
            assert np.array_equal(full[:, :8].copy().view(np.int64)[:, 0], np.arange(300))
    big = O.fill_rows_at(np.array([999_999, 1 << 33], dtype=np.int64), 16, 5)  # row numbers beyond any test buffer
    assert big[:, :8].copy().view(np.int64)[:, 0].tolist() == [999_999, 1 << 33]


def test_mappo_gae_golden_bit_exact(golden_dir):
    """the GAE block of MAPPOLearner.learn (mappo.py:478-500), run by the reference's own statements
    (tests/golden/make_golden.py::make_mappo): steps 1..25 of the [26, B, 1] arrays with V_26 = critic(last obs).
    The reference's discounted_r is our `ret` bit for bit; its advantage is (discounted_r - V), formed afterwards."""
    g = np.load(os.path.join(golden_dir, "mappo_gae_ref.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 4
    for n in names:
        reward, value, mask = g[f"{n}/reward"][..., 0], g[f"{n}/value"][..., 0], g[f"{n}/mask"][..., 0]
        last = g[f"{n}/last_value"][..., 0]
        adv, ret = O.gae(reward[1:], value[1:], None, last, 1.0 - mask[1:], float(g[f"{n}/gamma"]), float(g[f"{n}/lam"]), layout=0)
        dr = g[f"{n}/discounted_r"][..., 0]
        assert np.array_equal(ret.view(np.uint32), dr[1:].view(np.uint32))
        assert np.all(dr[0] == 0)  # never written by the reference loop
        assert np.array_equal((ret - value[1:]).view(np.uint32), g[f"{n}/advantage"][..., 0].view(np.uint32))
        assert np.allclose(adv, g[f"{n}/advantage"][..., 0], rtol=0, atol=2e-6)  # (gae + V) - V against gae itself


def test_coma_td_lambda_golden_bit_exact(golden_dir):
    """COMALearner.build_td_lambda_targets (coma.py:489-501) run by the reference's own method
    (make_golden.py::make_coma) against the oracle's restatement"""
    g = np.load(os.path.join(golden_dir, "coma_td_lambda_ref.npz"))
    names = sorted({k.split("/")[0] for k in g.files})
    assert len(names) == 4
    for n in names:
        got = O.td_lambda_targets(g[f"{n}/rewards"], g[f"{n}/terminated"], g[f"{n}/mask"], g[f"{n}/target_qs"],
                                  float(g[f"{n}/gamma"]), float(g[f"{n}/td_lambda"]), int(g[f"{n}/max_t"]))
        assert np.array_equal(got.view(np.uint32), g[f"{n}/ret"].view(np.uint32)), n


def test_vtrace_golden_through_the_whole_oracle(golden_dir):
    """vtrace.py:79-105 run by the reference's own module against the oracle's restatement of all of it (exp, clips,
    deltas, recurrence, vs, pg advantages); numpy's exp and libm's expf may differ in the last bit"""
    g = np.load(os.path.join(golden_dir, "vtrace_ref.npz"))
    for n in sorted({k.split("/")[0] for k in g.files}):
        args = [g[f"{n}/{k}"] for k in ("log_rhos", "discounts", "rewards", "values", "boot")]
        vs, pg = O.vtrace(*args, 1.0, 1.0, 1.0, g[f"{n}/masks"])
        for got, want in ((vs, g[f"{n}/vs"]), (pg, g[f"{n}/pg"])):
            assert np.abs(got - want).max() <= 1e-5 * max(1.0, float(np.abs(want).max()))


def test_priority_pow_and_weights_against_an_independent_float64_restatement():
    """include/mrl_arith.h (mrl_detpowf, mrl_priority_pow, mrl_is_weight) is compiled into BOTH the
    CUDA library and this oracle, so GPU-vs-oracle parity of p**alpha and of the importance weights compares a
    source with itself.  This test restates both from the published formulas in numpy float64 WITHOUT that
    header -- PER (Schaul et al. 2015, cited at priority_replay_buffer.py:33):
        leaf = priority ** alpha;  P(i) = leaf_i / sum;  w_i = (N * P(i)) ** -beta / max_j w_j,
        max_j w_j = (N * min_leaf / sum) ** -beta
    -- and checks the oracle's tree leaves and weights against it: leaves equal the correctly rounded float32 of
    the float64 power (the header's contract), weights agree to float32 rounding of the few fp32 steps."""
    rng = np.random.default_rng(123)
    cap, batch, alpha, beta = 4096, 256, 0.6, 0.4
    o = O.PriorityBufferOracle(alpha, cap, batch, [(1,)], [np.int32], 1, 2)
    o.insert_batch([np.arange(cap, dtype=np.int32)[:, None]])
    pr = (np.abs(rng.normal(size=cap)) * 5 + 1e-3).astype(np.float32)
    o.update_priorities(np.arange(cap), pr)
    leaves64 = np.power(pr.astype(np.float64), np.float64(np.float32(alpha)))
    s, m = o.tree()
    assert np.array_equal(s[cap:2 * cap], leaves64.astype(np.float32))       # correctly rounded pow, every leaf
    assert abs(float(s[1]) - leaves64.sum()) <= 2e-6 * leaves64.sum()          # fp32 pairwise tree sum
    assert float(m[1]) == float(leaves64.astype(np.float32).min())
    u = rng.random(batch).astype(np.float32)
    idx, w, _ = o.sample(beta, uniforms=u)
    # stratified proportional draw, restated: stratum i covers mass [(i)/n, (i+1)/n) of the total
    cum = np.cumsum(leaves64)
    mass = (u.astype(np.float64) + np.arange(batch)) * (cum[-1] / batch)
    want_idx = np.searchsorted(cum, mass, side="left")
    agree = (want_idx == idx).mean()
    assert agree >= 0.99, agree                                                # fp32 prefix sums move a boundary draw by one leaf
    prob = leaves64[idx] / cum[-1]
    w64 = np.power(prob * cap, -np.float64(np.float32(beta))) / np.power(leaves64.min() / cum[-1] * cap, -np.float64(np.float32(beta)))
    assert np.allclose(w, w64, rtol=3e-6, atol=0)
    assert w.max() <= 1.0 + 1e-6


def test_exact_shard_weights_undo_the_sampling_probability():
    """with shards of unequal priority mass the default sharded weights (global sum) do not cancel the
    probability a sample is really drawn with, p_i / (world * S_rank).  With shard_weights = 1 they do: at beta = 1,
    w_i * P(i) is one constant over all samples of all ranks (the definition of an importance weight); with the
    default it differs between the ranks by the ratio of their masses."""
    rng = np.random.default_rng(17)
    world, cap, batch = 2, 256, 64
    bufs, stats = [], []
    for r in range(world):
        b = O.PriorityBufferOracle(1.0, cap, batch, [(1,)], [np.int32], 1, r)
        b.insert_batch([np.arange(cap, dtype=np.int32)[:, None]])
        b.update_priorities(np.arange(cap), (rng.random(cap) * (1 + 4 * r) + 0.05).astype(np.float32))  # rank 1: ~5x the mass
        bufs.append(b)
        stats.append(b.root_stats())
    g = np.stack(stats)
    u = rng.random(batch).astype(np.float32)
    try:
        spread = {}
        for exact in (0, 1):
            O.set_shard_weights(exact)
            prod = []
            for r, b in enumerate(bufs):
                idx, w = b.sample(1.0, uniforms=u, gather=False, gstats=g)
                p = b.tree()[0][cap + idx].astype(np.float64)
                prod.append(w.astype(np.float64) * p / (world * float(stats[r][0])))
                assert w.max() <= 1.0 + 1e-6
            allp = np.concatenate(prod)
            spread[exact] = allp.max() / allp.min()
        assert spread[1] < 1.0 + 1e-5, spread
        assert spread[0] > 2.0, spread  # the masses differ ~5x
    finally:
        O.set_shard_weights(0)
