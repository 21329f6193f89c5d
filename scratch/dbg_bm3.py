import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mindrl_b200 as M
from mindrl_b200 import _ffi
_ffi.set_option("bm_kernel", 2)
stages, B, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
_ffi.set_option("row_stages", stages)

t = torch.arange(T, device="cuda", dtype=torch.float32)
r = (t[None, :] + 100000.0 * torch.arange(B, device="cuda", dtype=torch.float32)[:, None]).contiguous()
v = torch.zeros((B, T), device="cuda"); nv = torch.zeros((B, T), device="cuda")
for rep in range(3):
    adv, ret = M.gae(r, v, 0.0, 0.0, next_value=nv, layout="batch_major")
    torch.cuda.synchronize()
    bad = (adv != r).nonzero().cpu().numpy()
    if len(bad):
        a = adv.cpu().numpy()
        print("rep", rep, "nbad", len(bad))
        last = None
        for b, tt in bad[:0]:
            src = a[b, tt]
            key = (b, tt // 256, int(src) - tt)
            if key != last:
                print("  row", b, "t", tt, "got", src, "delta", int(src) - int(r[b, tt].item()))
                last = key
print("done")
