import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindrl_b200 as M
from mindrl_b200 import _ffi
from oracle import oracle as O
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
_ffi.set_option("bm_kernel", 2)
def ranges(bad):
    out = []; start = None; prev = None
    for t in bad:
        if start is None: start = prev = t
        elif t == prev + 1: prev = t
        else: out.append((start, prev)); start = prev = t
    if start is not None: out.append((start, prev))
    return out
for B, T in [(3, 6160), (2, 8192), (1, 4096)]:
    rng = np.random.default_rng(B * 1009 + T)
    r = rng.normal(size=(B, T)).astype(np.float32); v = rng.normal(size=(B, T)).astype(np.float32)
    nv = rng.normal(size=(B, T)).astype(np.float32); lv = rng.normal(size=(B,)).astype(np.float32)
    for stages in (2, 3, 4):
        _ffi.set_option("row_stages", stages)
        for name, kw in (("nv", dict(next_value=dev(nv))), ("last", dict(last_value=dev(lv)))):
            adv_r, ret_r = O.gae(r, v, nv if name == "nv" else None, None if name == "nv" else lv, None, 0.99, 0.95, layout=1)
            for rep in range(3):
                adv, ret = M.gae(dev(r), dev(v), 0.99, 0.95, layout="batch_major", **kw)
                err = np.abs(adv.cpu().numpy() - adv_r)
                for b in range(B):
                    bad = np.nonzero(err[b] > 2e-4)[0]
                    if len(bad): print(B, T, "stages", stages, name, "rep", rep, "row", b, ranges(bad.tolist())[:6], flush=True)
        ref = O.discounted_return(r, None, lv, 0.99, layout=1)
        out = M.discounted_reward(dev(r), 0.99, last_value=dev(lv))
        err = np.abs(out.cpu().numpy() - ref)
        for b in range(B):
            bad = np.nonzero(err[b] > 2e-4 * np.abs(ref).max())[0]
            if len(bad): print(B, T, "stages", stages, "DR row", b, ranges(bad.tolist())[:6], flush=True)
print("done")
