"""Generates tests/golden/discounted_return_ref.npz by running the REFERENCE's own Python loop
(/root/reference/mindspore_rl/utils/discounted_return.py:84-92, the CPU path of DiscountedReturn).

MindSpore itself cannot be installed here, so the reference file is executed against a minimal
numpy stand-in for the handful of MindSpore names it touches (Tensor, nn.Cell, P.ZerosLike,
context.get_context, rl_ops.DiscountedReturn).  The stand-in follows MindSpore's type promotion
(an integer tensor times a float32 tensor is float32).  The loop body, its operation order and
its indexing are the reference's, untouched: the file is loaded from /root/reference at run time
and never copied.

A second fixture, ppo_scans_ref.npz, comes from PPOLearner.learn's nested functions
`discounted_reward` and `gae` (/root/reference/mindspore_rl/algorithm/ppo/ppo.py:320-350).  They
cannot be imported (they are closures inside learn()), so their FunctionDef nodes are taken from
the reference file with `ast` at run time and compiled unmodified against a stand-in `self`
(add, mul, reshape, squeeze, zeros_like, zero_int, critic_net = a lookup of precomputed values).

mappo_gae_ref.npz: the GAE block of MAPPOLearner.learn (algorithm/mappo/mappo.py:478-500), inline statements
taken out of the method with `ast`; coma_td_lambda_ref.npz: COMALearner.build_td_lambda_targets
(algorithm/coma/coma.py:489-501), the method itself; vtrace_ref.npz: the reference's vtrace.py imported whole.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/mindspore_rl/utils/discounted_return.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "discounted_return_ref.npz")


class Tensor(np.ndarray):
    """ndarray with MindSpore-like promotion: float32 wins over integer operands"""

    def __new__(cls, data, dtype=None):
        return np.asarray(data, dtype=dtype).view(cls)

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        raw = [np.asarray(i).view(np.ndarray) if isinstance(i, np.ndarray) else i for i in inputs]
        kwargs.pop("out", None)  # `x += 1` builds a new tensor, as MindSpore's tensors do (never mutates a shared one)
        out = getattr(ufunc, method)(*raw, **kwargs)
        had64 = any(isinstance(i, np.ndarray) and i.dtype == np.float64 for i in inputs)
        if isinstance(out, np.ndarray) and out.dtype == np.float64 and not had64:
            out = out.astype(np.float32)
        return np.asarray(out).view(Tensor)

    def asnumpy(self):
        return np.asarray(self).view(np.ndarray)


def install_shim():
    ms = types.ModuleType("mindspore")
    ms.float32, ms.float16, ms.bool_, ms.int32 = np.float32, np.float16, np.bool_, np.int32
    ms.Tensor = Tensor

    class Cell:
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return self.construct(*a, **k)

    nn = types.ModuleType("mindspore.nn")
    nn.Cell = Cell
    ms.nn = nn
    context = types.ModuleType("mindspore.context")
    context.get_context = lambda key: "CPU"
    ms.context = context
    ops = types.ModuleType("mindspore.ops")
    operations = types.ModuleType("mindspore.ops.operations")
    operations.ZerosLike = lambda: (lambda x: Tensor(np.zeros_like(np.asarray(x))))
    rl = types.ModuleType("mindspore.ops.operations._rl_inner_ops")
    rl.DiscountedReturn = lambda gamma: None
    operations._rl_inner_ops = rl
    ops.operations = operations
    ms.ops = ops
    for name, mod in [("mindspore", ms), ("mindspore.nn", nn), ("mindspore.context", context),
                      ("mindspore.ops", ops), ("mindspore.ops.operations", operations),
                      ("mindspore.ops.operations._rl_inner_ops", rl)]:
        sys.modules[name] = mod


def load_reference():
    install_shim()
    spec = importlib.util.spec_from_file_location("ref_discounted_return", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DiscountedReturn


PPO_REF = "/root/reference/mindspore_rl/algorithm/ppo/ppo.py"
OUT_PPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppo_scans_ref.npz")


def load_ppo_closures(next_values):
    """compiles ppo.py's nested discounted_reward()/gae() as they stand in the reference"""
    import ast
    tree = ast.parse(open(PPO_REF).read())
    wanted = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in ("discounted_reward", "gae"):
            wanted[node.name] = node

    class Self:
        zero_int = 0
        add = staticmethod(lambda a, b: a + b)
        mul = staticmethod(lambda a, b: a * b)
        reshape = staticmethod(lambda a, shp: np.asarray(a).reshape(shp).view(Tensor))
        squeeze = staticmethod(lambda a: np.squeeze(np.asarray(a), -1).view(Tensor))
        zeros_like = staticmethod(lambda a: Tensor(np.zeros_like(np.asarray(a))))
        critic_net = staticmethod(lambda x: next_values)

    ns = {"self": Self(), "len": len}
    mod = ast.Module(body=[wanted["discounted_reward"], wanted["gae"]], type_ignores=[])
    exec(compile(mod, PPO_REF, "exec"), ns)
    return ns["discounted_reward"], ns["gae"]


def make_ppo():
    install_shim()
    rng = np.random.default_rng(7)
    cases = {}
    for name, (B, T), gamma, lam in [("a", (3, 17), 0.99, 0.95), ("b", (30, 200), 0.99, 0.95),
                                     ("c", (5, 128), 0.9, 0.8), ("d", (2, 1), 0.99, 0.95),
                                     ("e", (7, 260), 0.995, 0.97)]:
        reward = rng.normal(size=(B, T, 1)).astype(np.float32)
        value = rng.normal(size=(B, T, 1)).astype(np.float32)
        next_value = rng.normal(size=(B, T, 1)).astype(np.float32)
        disc, gae = load_ppo_closures(Tensor(next_value))
        g = Tensor([gamma], np.float32).reshape(())  # self.gamma = Tensor(gamma, float32), ppo.py
        v_last = Tensor(np.zeros((B,), np.float32))
        dr = disc(Tensor(np.squeeze(reward, -1)), v_last, g)
        adv = gae(Tensor(reward), None, Tensor(value), v_last, g, lam)
        cases[f"{name}/reward"] = reward[..., 0]
        cases[f"{name}/value"] = value[..., 0]
        cases[f"{name}/next_value"] = next_value[..., 0]
        cases[f"{name}/gamma"] = np.float32(gamma)
        cases[f"{name}/lam"] = np.float32(lam)
        cases[f"{name}/discounted_r"] = np.asarray(dr).view(np.ndarray).astype(np.float32)
        cases[f"{name}/advantage"] = np.asarray(adv).view(np.ndarray).astype(np.float32)
        assert np.asarray(adv).dtype == np.float32 and np.asarray(dr).dtype == np.float32
    np.savez_compressed(OUT_PPO, **cases)
    print("wrote", OUT_PPO, os.path.getsize(OUT_PPO), "bytes")


VTRACE_REF = "/root/reference/mindspore_rl/algorithm/impala/vtrace.py"
OUT_VTRACE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vtrace_ref.npz")


def make_vtrace():
    """runs the reference's get_vs_and_advantages (vtrace.py:48-108) on the numpy stand-in"""
    install_shim()
    ms = sys.modules["mindspore"]
    ops = sys.modules["mindspore.ops"]
    w = lambda a: np.asarray(a).view(Tensor)  # noqa: E731
    ops.exp = lambda x: w(np.exp(np.asarray(x, dtype=np.float32)))
    ops.minimum = lambda a, b: w(np.minimum(np.asarray(a), np.asarray(b)))
    ops.concat = lambda xs, axis=0: w(np.concatenate([np.asarray(x) for x in xs], axis=axis))
    ops.expand_dims = lambda x, axis: w(np.expand_dims(np.asarray(x), axis))
    ops.zeros_like = lambda x: w(np.zeros_like(np.asarray(x)))
    ops.stack = lambda xs: w(np.stack([np.asarray(x) for x in xs]))
    ops.flip = lambda x, dims: w(np.flip(np.asarray(x), axis=tuple(dims)))
    ops.stop_gradient = lambda x: x
    ops.transpose = lambda x, perm: w(np.transpose(np.asarray(x), perm))
    ms.ops = ops
    spec = importlib.util.spec_from_file_location("ref_vtrace", VTRACE_REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(11)
    cases = {}
    for name, (T, B) in [("a", (5, 3)), ("b", (20, 16)), ("c", (100, 8)), ("d", (1, 4))]:
        log_rhos = (rng.normal(size=(T, B)) * 0.5).astype(np.float32)
        discounts = np.full((T, B), 0.99, np.float32)
        rewards = rng.normal(size=(T, B)).astype(np.float32)
        values = rng.normal(size=(T, B)).astype(np.float32)
        boot = rng.normal(size=(B,)).astype(np.float32)
        masks = (rng.random((T, B)) > 0.1).astype(np.float32)
        vs, pg = mod.get_vs_and_advantages(Tensor(log_rhos), Tensor(discounts), Tensor(rewards), Tensor(values),
                                           Tensor(boot), 1.0, 1.0, 1.0, Tensor(masks))
        for k, v in dict(log_rhos=log_rhos, discounts=discounts, rewards=rewards, values=values, boot=boot,
                         masks=masks, vs=np.asarray(vs).view(np.ndarray).astype(np.float32),
                         pg=np.asarray(pg).view(np.ndarray).astype(np.float32)).items():
            cases[f"{name}/{k}"] = v
        assert np.asarray(vs).dtype == np.float32
    np.savez_compressed(OUT_VTRACE, **cases)
    print("wrote", OUT_VTRACE, os.path.getsize(OUT_VTRACE), "bytes")


MAPPO_REF = "/root/reference/mindspore_rl/algorithm/mappo/mappo.py"
OUT_MAPPO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mappo_gae_ref.npz")


def make_mappo():
    """MAPPOLearner.learn's GAE block (mappo.py:478-500) is inline code, not a function: its statements -- from
    `gae = self.zeros_like(last_value)` to `advantage = (...)[1:]` -- are taken out of learn()'s FunctionDef with
    `ast` at run time and compiled UNMODIFIED into a function of the names they read (self, reward, value, mask,
    last_value, last_episode_mask, last_value_prediction, ms).  self.gamma is a float32 tensor and
    self.td_lambda a Python float, as at mappo.py:430-431; value_normalizer.denormalize is the identity (the
    scan takes de-normalised values)."""
    import ast
    install_shim()
    ms = sys.modules["mindspore"]
    src = open(MAPPO_REF).read()
    tree = ast.parse(src)
    learn = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "learn"]
    assert len(learn) == 1
    body = learn[0].body
    first = next(i for i, st in enumerate(body) if isinstance(st, ast.Assign) and ast.unparse(st).startswith("gae = self.zeros_like(last_value)"))
    last = next(i for i, st in enumerate(body) if isinstance(st, ast.Assign) and ast.unparse(st).startswith("advantage = "))
    assert 478 <= body[first].lineno and body[last].end_lineno <= 500, (body[first].lineno, body[last].end_lineno)
    fn = ast.FunctionDef(
        name="mappo_gae_block",
        args=ast.arguments(posonlyargs=[], args=[ast.arg(arg=a) for a in
                                                 ("self", "reward", "value", "mask", "last_value", "last_episode_mask",
                                                  "last_value_prediction", "ms")],
                           kwonlyargs=[], kw_defaults=[], defaults=[]),
        body=body[first:last + 1] + [ast.Return(value=ast.Tuple(elts=[ast.Name(id="discounted_r", ctx=ast.Load()),
                                                                     ast.Name(id="advantage", ctx=ast.Load())], ctx=ast.Load()))],
        decorator_list=[], type_params=[])
    mod = ast.fix_missing_locations(ast.Module(body=[fn], type_ignores=[]))
    ns = {}
    exec(compile(mod, MAPPO_REF, "exec"), ns)

    class Norm:
        denormalize = staticmethod(lambda x: x)

    class Self:
        zero = Tensor(0, np.float32)
        zeros_like = staticmethod(lambda a: Tensor(np.zeros_like(np.asarray(a))))
        value_normalizer = Norm()

    rng = np.random.default_rng(31)
    cases = {}
    for name, B, gamma, lam in [("a", 1, 0.99, 0.95), ("b", 8, 0.99, 0.95), ("c", 64, 0.9, 0.8), ("d", 33, 0.995, 0.97)]:
        me = Self()
        me.gamma = Tensor(gamma, np.float32)
        me.td_lambda = lam
        reward = rng.normal(size=(26, B, 1)).astype(np.float32)
        value = rng.normal(size=(26, B, 1)).astype(np.float32)
        mask = (rng.random((26, B, 1)) > 0.15).astype(np.float32)
        last_value = rng.normal(size=(B, 1)).astype(np.float32)
        dr, adv = ns["mappo_gae_block"](me, Tensor(reward), Tensor(value), Tensor(mask), Tensor(last_value),
                                        Tensor(mask[-1]), Tensor(value[-1]), ms)
        dr, adv = np.asarray(dr).view(np.ndarray), np.asarray(adv).view(np.ndarray)
        assert dr.dtype == np.float32 and adv.dtype == np.float32 and dr.shape == (26, B, 1) and adv.shape == (25, B, 1)
        for k, v in dict(reward=reward, value=value, mask=mask, last_value=last_value, gamma=np.float32(gamma),
                         lam=np.float32(lam), discounted_r=dr, advantage=adv).items():
            cases[f"{name}/{k}"] = v
    np.savez_compressed(OUT_MAPPO, **cases)
    print("wrote", OUT_MAPPO, os.path.getsize(OUT_MAPPO), "bytes")


COMA_REF = "/root/reference/mindspore_rl/algorithm/coma/coma.py"
OUT_COMA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "coma_td_lambda_ref.npz")


def make_coma():
    """COMALearner.build_td_lambda_targets (coma.py:489-501), its FunctionDef taken from the reference file with
    `ast` and compiled unmodified; F.zeros_like is the only outside name it reads.  gamma and td_lambda are
    Python floats (coma.py:395-396)."""
    import ast
    install_shim()
    tree = ast.parse(open(COMA_REF).read())
    fns = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "build_td_lambda_targets"]
    assert len(fns) == 1 and 489 <= fns[0].lineno <= 490
    mod = ast.Module(body=fns, type_ignores=[])

    class F:
        zeros_like = staticmethod(lambda a: Tensor(np.zeros_like(np.asarray(a))))

    ns = {"F": F}
    exec(compile(mod, COMA_REF, "exec"), ns)
    rng = np.random.default_rng(47)
    cases = {}
    for name, (B, T, A), gamma, lam, max_t in [("a", (2, 5, 3), 0.99, 0.8, 5), ("b", (32, 60, 5), 0.99, 0.8, 60),
                                               ("c", (8, 100, 2), 0.95, 0.6, 73), ("d", (4, 1, 1), 0.99, 0.8, 1)]:
        rewards = rng.normal(size=(B, T, 1)).astype(np.float32)
        terminated = (rng.random((B, T, 1)) < 0.05).astype(np.float32)
        mask = (rng.random((B, T, 1)) > 0.1).astype(np.float32)
        target_qs = rng.normal(size=(B, T + 1, A)).astype(np.float32)
        ret = ns["build_td_lambda_targets"](None, Tensor(rewards), Tensor(terminated), Tensor(mask), Tensor(target_qs),
                                            gamma, lam, max_t)
        ret = np.asarray(ret).view(np.ndarray)
        assert ret.dtype == np.float32 and ret.shape == target_qs.shape
        for k, v in dict(rewards=rewards, terminated=terminated, mask=mask, target_qs=target_qs, gamma=np.float64(gamma),
                         td_lambda=np.float64(lam), max_t=np.int64(max_t), ret=ret).items():
            cases[f"{name}/{k}"] = v
    np.savez_compressed(OUT_COMA, **cases)
    print("wrote", OUT_COMA, os.path.getsize(OUT_COMA), "bytes")


def main():
    make_mappo()
    make_coma()
    make_vtrace()
    make_ppo()
    DR = load_reference()
    rng = np.random.default_rng(20261019)
    cases = {}
    # the reference's own two known-answer vectors (tests/st/test_discounted_return.py:39-53)
    specs = [("kat0", (4, 1), (4, 1), (1,), 0.99), ("kat1", (4,), (4,), (), 0.99)]
    shapes = [((7, 3), (7, 3), (3,)), ((64, 8), (64, 8), (8,)), ((33, 5, 4), (33, 5, 1), (5, 4)),
              ((200, 2), (200, 2), (2,)), ((1, 6), (1, 6), (6,)), ((300, 16, 2), (300, 16, 1), (16, 2))]
    for i, (rs, ds, ls) in enumerate(shapes):
        for gamma in (0.99, 0.9, 1.0, 0.0):
            specs.append((f"rand{i}_g{gamma}", rs, ds, ls, gamma))
    for name, rs, ds, ls, gamma in specs:
        if name == "kat0":
            reward = np.ones(rs, np.float32)
            done = np.array([[False], [False], [True], [False]])
            last = np.array([2.0], np.float32)
        elif name == "kat1":
            reward = np.ones(rs, np.float32)
            done = np.array([False, True, False, True])
            last = np.array(2.0, np.float32)
        else:
            reward = rng.normal(size=rs).astype(np.float32)
            done = rng.random(size=ds) < 0.1
            last = rng.normal(size=ls).astype(np.float32)
        net = DR(gamma=gamma)
        out = net(Tensor(reward, np.float32), Tensor(done), Tensor(last, np.float32))
        out = np.asarray(out).view(np.ndarray)
        assert out.dtype == np.float32 and out.shape == reward.shape
        cases[name + "/reward"] = reward
        cases[name + "/done"] = done
        cases[name + "/last"] = last
        cases[name + "/gamma"] = np.float32(gamma)
        cases[name + "/out"] = out
    # sanity: the reference's own expectations
    assert np.allclose(cases["kat0/out"], np.array([[2.9701], [1.99], [1.0], [2.98]], np.float32))
    assert np.allclose(cases["kat1/out"], np.array([1.99, 1, 1.99, 1.0], np.float32))
    np.savez_compressed(OUT, **cases)
    print("wrote", OUT, len(specs), "cases", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
