"""one launch of each scan kernel of the C4 shape (for ncu): GAE and DiscountedReturn in both layouts, GAE + normalisation"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
T, B = 2048, 4096
g = torch.Generator(device="cuda"); g.manual_seed(4)
r = torch.randn((T, B), device="cuda", generator=g)
v = torch.randn((T + 1, B), device="cuda", generator=g)
d = (torch.rand((T, B), device="cuda", generator=g) < 1 / 200).to(torch.uint8)
adv = torch.empty((T, B), device="cuda"); ret = torch.empty((T, B), device="cuda")
p = lambda x: x.data_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):  # default kernels of this shape (TMA strips / rows)
    flush.fill_(i)
    _ffi.check(L.mrl_gae(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), T, B, 0.99, 0.95, 0, 0, 0))
    flush.fill_(i)
    _ffi.check(L.mrl_gae(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), T, B, 0.99, 0.95, 1, 0, 0))
    flush.fill_(i)
    _ffi.check(L.mrl_discounted_return(p(r), p(d), p(v[T]), p(adv), T, B, 1, 0.99, 0, 0, 0))
    flush.fill_(i)
    _ffi.check(L.mrl_discounted_return(p(r), p(d), p(v[T]), p(adv), T, B, 1, 0.99, 1, 0, 0))
    flush.fill_(i)
    _ffi.check(L.mrl_gae_normalized(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), T, B, 0.99, 0.95, 1e-8, 0, 0, 0))
torch.cuda.synchronize()
print("ok")
