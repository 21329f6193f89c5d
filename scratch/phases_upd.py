import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
cap, batch = 1_000_000, 256
rows = [16, 4, 4, 16]
rb = np.asarray(rows, dtype=np.int64)
h = C.c_int64(-1)
_ffi.check(L.mrl_prb_create(cap, 0.6, 4, rb.ctypes.data, 7, 0, C.byref(h)))
cols = (C.c_void_p * 4)()
_ffi.check(L.mrl_prb_columns(h.value, cols))
_ffi.check(L.mrl_prb_push_batch(h.value, cap, cols, 0))
prio = torch.randn((8, batch), device="cuda").abs_() + 1e-3
marks = torch.zeros(32, dtype=torch.int64, device="cuda")
_ffi.set_option("debug_timers", marks.data_ptr())
acc = np.zeros(12); cnt = 0
for i in range(200):
    idx = torch.randint(0, cap, (batch,), dtype=torch.int64, device="cuda")
    idx[0] = 5  # block 0 owns subtree 0: make sure it has an entry
    _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0))
    torch.cuda.synchronize()
    m = marks.cpu().numpy().astype(np.float64)
    seq = m[[18, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29]]
    if i >= 5 and np.all(np.diff(seq) >= 0):
        acc[:10] += np.diff(seq) / 1.965e3; cnt += 1
print("block 0, us: scan (start->warp path) %.2f | idx + issue sibling loads %.2f | pow %.2f | leaf arbitration %.2f | walk %.2f | last store %.2f | -> after if %.2f | max reduce + atomicMax %.2f | threadfence %.2f | sync + ticket %.2f  (n=%d)" % (*(acc[:10] / cnt), cnt))
