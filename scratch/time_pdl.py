"""C2 step (sample + update_priorities) with and without programmatic dependent launch"""
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
def step(i):
    _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, 0))
    _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0))
for pdl in (1, 0, 1, 0):
    _ffi.set_option("pdl", pdl)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    # (a) flushed, events per step
    tot = 0.0
    for i in range(100):
        flush.fill_(i & 255); fsrc.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(i); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    # (b) 200 steps back to back in one event pair (warm tree)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): step(i)
    e1.record(); torch.cuda.synchronize()
    print("pdl", pdl, "flushed us/step", round(tot * 10, 2), "back-to-back us/step", round(e0.elapsed_time(e1) * 5, 2), flush=True)
