"""one launch of each TMA-ring scan kernel at C4 size (for ncu)"""
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
cols = int(sys.argv[1]) if len(sys.argv) > 1 else 16
_ffi.set_option("tm_kernel", 6); _ffi.set_option("strip_cols", cols); _ffi.set_option("bm_kernel", 2)
_ffi.set_option("strip_rpu", int(sys.argv[2]) if len(sys.argv) > 2 else 16)
for _ in range(2):
    _ffi.check(L.mrl_gae(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), T, B, 0.99, 0.95, 0, 0, 0))
    _ffi.check(L.mrl_discounted_return(p(r), p(d), p(v[T]), p(adv), T, B, 1, 0.99, 0, 0, 0))
    _ffi.check(L.mrl_gae(p(r), p(v), None, p(v[T]), p(d), p(adv), p(ret), 2048, 4096, 0.99, 0.95, 1, 0, 0))
    _ffi.check(L.mrl_discounted_return(p(r), p(d), p(v[T]), p(adv), 2048, 4096, 1, 0.99, 1, 0, 0))
torch.cuda.synchronize()
print("ok")
