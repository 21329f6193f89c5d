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
idx = torch.empty(batch, dtype=torch.int64, device="cuda")
w = torch.empty(batch, dtype=torch.float32, device="cuda")
outs = [torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows]
tab = _ffi.ptr_table([o.data_ptr() for o in outs])
h_idx = torch.empty(batch, dtype=torch.int64).pin_memory()
h_w = torch.empty(batch, dtype=torch.float32).pin_memory()
h_out = [torch.empty((batch, r), dtype=torch.uint8).pin_memory() for r in rows]
h_tab = _ffi.ptr_table([o.data_ptr() for o in h_out])
prio = (torch.randn((8, batch)).abs_() + 1e-3).pin_memory()
hv = h.value
N = 2000
for rep in range(2):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(N):
        L.mrl_prb_sample_host(hv, batch, 0.4, None, h_idx.data_ptr(), h_w.data_ptr(), h_tab, 0)
    t1 = time.perf_counter()
    out = (C.c_double * 8)()
    L.mrl_debug_host_times(out)
    print("sample_host us/call", round((t1 - t0) / N * 1e6, 2), "inside: wait+ensure %.2f launch %.2f spin %.2f" % (out[0], out[1], out[2]))
    t0 = time.perf_counter()
    for i in range(N):
        L.mrl_prb_sample(hv, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, 0)
        L.mrl_stream_sync(0)
    t1 = time.perf_counter()
    print("sample (device out) + stream_sync us/call", round((t1 - t0) / N * 1e6, 2))
    t0 = time.perf_counter()
    for i in range(N):
        L.mrl_prb_sample(hv, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), None, 0)
        L.mrl_stream_sync(0)
    t1 = time.perf_counter()
    print("sample (no rows) + stream_sync us/call", round((t1 - t0) / N * 1e6, 2))
    t0 = time.perf_counter()
    for i in range(N):
        L.mrl_prb_update_host(hv, h_idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0)
        L.mrl_stream_sync(0)
    t1 = time.perf_counter()
    print("update_host + stream_sync us/call", round((t1 - t0) / N * 1e6, 2))
    t0 = time.perf_counter()
    for i in range(N):
        L.mrl_prb_update_host(hv, h_idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0)
    t1 = time.perf_counter()
    L.mrl_stream_sync(0)
    print("update_host back to back (no sync) us/call", round((t1 - t0) / N * 1e6, 2))
    t0 = time.perf_counter()
    for i in range(N):
        L.mrl_launch_count()
    t1 = time.perf_counter()
    print("ctypes null call us", round((t1 - t0) / N * 1e6, 2))
    t0 = time.perf_counter()
    for i in range(N):
        h_idx.data_ptr()
    t1 = time.perf_counter()
    print("data_ptr() us", round((t1 - t0) / N * 1e6, 2))
