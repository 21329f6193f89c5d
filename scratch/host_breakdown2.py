import ctypes as C, os, sys, time
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
idx = torch.randint(0, cap, (batch,), dtype=torch.int64, device="cuda")
h_idx = idx.cpu().pin_memory()
prio = (torch.randn((8, batch)).abs_() + 1e-3).pin_memory()
d_prio = prio.cuda()
hv = h.value
N = 2000
for inline in (1, 0, 1, 0):
    _ffi.set_option("inline_update", inline)
    tc = ts = 0.0
    for i in range(N):
        t0 = time.perf_counter()
        L.mrl_prb_update_host(hv, h_idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0)
        t1 = time.perf_counter()
        L.mrl_stream_sync(0)
        t2 = time.perf_counter()
        tc += t1 - t0; ts += t2 - t1
    print("inline", inline, "update_host call us", round(tc / N * 1e6, 2), "then sync us", round(ts / N * 1e6, 2))
tc = ts = 0.0
for i in range(N):
    t0 = time.perf_counter()
    L.mrl_prb_update(hv, idx.data_ptr(), d_prio[i % 8].data_ptr(), batch, 0)
    t1 = time.perf_counter()
    L.mrl_stream_sync(0)
    t2 = time.perf_counter()
    tc += t1 - t0; ts += t2 - t1
print("device update call us", round(tc / N * 1e6, 2), "then sync us", round(ts / N * 1e6, 2))
