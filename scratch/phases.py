"""phase marks inside the PER kernels (cold L2 after a flush vs warm), tree prefetch on / off"""
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
fsrc = torch.zeros(64 << 20, device="cuda")
marks = torch.zeros(32, dtype=torch.int64, device="cuda")
_ffi.set_option("debug_timers", marks.data_ptr())
def run(cold, pf, n=50):
    _ffi.set_option("tree_prefetch", pf)
    acc = np.zeros(32)
    for i in range(n):
        if cold:
            flush.fill_(i & 255); fsrc.sum()
        _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, 0))
        torch.cuda.synchronize()
        m = marks.cpu().numpy().astype(np.float64)
        s = m[[0, 1, 3, 4, 5]]
        if cold:
            flush.fill_(i & 255); fsrc.sum()
        _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0))
        torch.cuda.synchronize()
        m = marks.cpu().numpy().astype(np.float64)
        u = m[8:17]
        acc[:4] += np.diff(s) / 1.965e3  # SM cycles -> us at 1965 MHz
        acc[8:16] += np.diff(u) / 1e3     # ns -> us
    acc /= n
    print("cold" if cold else "warm", "prefetch", pf,
          "| sample (CTA 0, us): root+mass %.2f descent %.2f weights %.2f rows %.2f" % tuple(acc[:4]),
          "| update (us): scan %.2f (9-10) %.2f (10-11) %.2f ticket %.2f (12-13) %.2f load %.2f (14-16) %.2f" %
          (acc[8], acc[9], acc[10], acc[11], acc[12], acc[13], acc[14] + acc[15]), flush=True)
for cold in (1, 0):
    for pf in (1, 0):
        run(cold, pf)
