import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
T, B = 2048, 4096
g = torch.Generator(device="cuda"); g.manual_seed(1)
r = torch.randn((T, B), device="cuda", generator=g); v = torch.randn((T + 1, B), device="cuda", generator=g)
d = (torch.rand((T, B), device="cuda", generator=g) < 0.005).to(torch.uint8)
o = torch.empty((T, B), device="cuda"); o2 = torch.empty((T, B), device="cuda")
marks = torch.zeros(32, dtype=torch.int64, device="cuda")
_ffi.set_option("debug_timers", marks.data_ptr())
P = lambda t: t.data_ptr()
names = ["wait full", "LDS + release", "maps + compose", "barrier", "chain", "wait_read", "replay + STS", "fence + store"]
for rpu in (16, 32):
    for groups in (1, 2):
        if rpu == 32 and groups == 2: continue
        _ffi.set_option("strip_rpu", rpu); _ffi.set_option("strip_groups", groups)
        for name, fn in (("GAE", lambda: _ffi.check(L.mrl_gae(P(r), P(v), None, P(v[T]), P(d), P(o), P(o2), T, B, 0.99, 0.95, 0, 0, 0))),
                         ("DR", lambda: _ffi.check(L.mrl_discounted_return(P(r), P(d), P(v[T]), P(o), T, B, 1, 0.99, 0, 0, 0)))):
            acc = np.zeros(8)
            for i in range(10):
                fn(); torch.cuda.synchronize()
                if i >= 2: acc += marks.cpu().numpy()[:8]
            acc /= 8
            tiles = T // (8 * rpu)
            print(f"{name} rpu {rpu} groups {groups}: per tile (us) " + ", ".join(f"{n} {a / tiles / 1965:.3f}" for n, a in zip(names, acc)) + f" | sum {acc.sum() / 1965:.2f} us over {tiles} tiles", flush=True)
