import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindrl_b200 as M
from mindrl_b200 import _ffi
T, B = int(sys.argv[1]), int(sys.argv[2])
_ffi.set_option("tm_kernel", 6)
r = torch.randn((T, B), device="cuda")
out = M.discounted_reward(r, 0.9, layout="time_major", mode="parallel")
torch.cuda.synchronize()
x = torch.zeros(B, device="cuda"); ref = torch.empty_like(r)
for t in range(T - 1, -1, -1):
    x = r[t] + 0.9 * x; ref[t] = x
print("maxerr", float((out - ref).abs().max()))
