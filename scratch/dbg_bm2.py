import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindrl_b200 as M
from mindrl_b200 import _ffi
from oracle import oracle as O
dev = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
_ffi.set_option("bm_kernel", 2)
which, stages, B, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
_ffi.set_option("row_stages", stages)
def ranges(bad):
    out = []; start = None; prev = None
    for t in bad:
        if start is None: start = prev = t
        elif t == prev + 1: prev = t
        else: out.append((start, prev)); start = prev = t
    if start is not None: out.append((start, prev))
    return out
rng = np.random.default_rng(B * 1009 + T)
r = rng.normal(size=(B, T)).astype(np.float32); v = rng.normal(size=(B, T)).astype(np.float32)
nv = rng.normal(size=(B, T)).astype(np.float32); lv = rng.normal(size=(B,)).astype(np.float32)
rd, vd, nvd, lvd = dev(r), dev(v), dev(nv), dev(lv)
torch.cuda.synchronize()
for rep in range(2):
    if which == "nv":
        ref, _ = O.gae(r, v, nv, None, None, 0.99, 0.95, layout=1); got, _ = M.gae(rd, vd, 0.99, 0.95, next_value=nvd, layout="batch_major")
    elif which == "last":
        ref, _ = O.gae(r, v, None, lv, None, 0.99, 0.95, layout=1); got, _ = M.gae(rd, vd, 0.99, 0.95, last_value=lvd, layout="batch_major")
    else:
        ref = O.discounted_return(r, None, lv, 0.99, layout=1); got = M.discounted_reward(rd, 0.99, last_value=lvd)
    err = np.abs(got.cpu().numpy() - ref)
    for b in range(B):
        bad = np.nonzero(err[b] > 2e-4 * max(1, np.abs(ref).max()))[0]
        if len(bad): print(which, stages, B, T, "rep", rep, "row", b, ranges(bad.tolist())[:6], flush=True)
print("ok", which, stages, B, T)
