"""a few sample+update steps on the C2 buffer (for ncu)"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
cap, batch = 1_000_000, int(sys.argv[2]) if len(sys.argv) > 2 else 256
rows = [16, 4, 4, 16]
rb = np.asarray(rows, dtype=np.int64)
h = C.c_int64(-1)
_ffi.check(L.mrl_prb_create(cap, 0.6, 4, rb.ctypes.data, 7, 0, C.byref(h)))
cols = (C.c_void_p * 4)()
_ffi.check(L.mrl_prb_columns(h.value, cols))
for c, r in enumerate(rows):
    _ffi.check(L.mrl_fill_rows(cols[c], cap, r, 1000 + c, 0))
_ffi.check(L.mrl_prb_push_batch(h.value, cap, cols, 0))
pr = torch.randn(cap, device="cuda").abs_() + 1e-3
idx_all = torch.arange(cap, device="cuda", dtype=torch.int64)
for s in range(0, cap, 8192):
    e = min(cap, s + 8192)
    _ffi.check(L.mrl_prb_update(h.value, idx_all[s:e].data_ptr(), pr[s:e].data_ptr(), e - s, 0))
idx = torch.empty(batch, dtype=torch.int64, device="cuda")
w = torch.empty(batch, dtype=torch.float32, device="cuda")
outs = [torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows]
tab = _ffi.ptr_table([o.data_ptr() for o in outs])
prio = torch.randn((8, batch), device="cuda").abs_() + 1e-3
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for i in range(reps):
    flush.fill_(i)
    _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, 0))
    _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0))
torch.cuda.synchronize()
print("ok")
