"""one launch of every scan kernel at C4 size (for ncu)"""
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
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(reps):
    _ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(), adv.data_ptr(), ret.data_ptr(), T, B, 0.99, 0.95, 0, 0, 0))
    _ffi.check(L.mrl_gae(r.data_ptr(), v.data_ptr(), None, v[T].data_ptr(), d.data_ptr(), adv.data_ptr(), ret.data_ptr(), B, T, 0.99, 0.95, 1, 0, 0))
    _ffi.check(L.mrl_discounted_return(r.data_ptr(), d.data_ptr(), v[T].data_ptr(), adv.data_ptr(), T, B, 1, 0.99, 0, 0, 0))
    _ffi.check(L.mrl_discounted_return(r.data_ptr(), d.data_ptr(), v[T].data_ptr(), adv.data_ptr(), B, T, 1, 0.99, 1, 0, 0))
torch.cuda.synchronize()
print("ok")
