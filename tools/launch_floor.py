"""what an (almost) empty kernel costs under the bench's timing discipline: after an L2 flush vs back to back"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
rb = np.asarray([16], dtype=np.int64)
h = C.c_int64(-1)
_ffi.check(L.mrl_prb_create(1024, 0.6, 1, rb.ctypes.data, 7, 0, C.byref(h)))
stats = torch.zeros(4, device="cuda")
x = torch.zeros(32, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); fsrc = torch.zeros(64 << 20, device="cuda")
def flushed(fn, n=100):
    tot = 0.0
    for i in range(n):
        flush.fill_(i & 255); fsrc.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / n * 1e3, 2)
def b2b(fn, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 2)
def noflush_single(fn, n=100):
    tot = 0.0
    for i in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / n * 1e3, 2)
for name, fn in [("nothing between the events", lambda: None),
                 ("torch x.add_(1), 32 floats", lambda: x.add_(1)),
                 ("mrl_prb_root_stats (1 thread, 3 loads)", lambda: L.mrl_prb_root_stats(h.value, stats.data_ptr(), 0)),
                 ("two of them", lambda: (L.mrl_prb_root_stats(h.value, stats.data_ptr(), 0), L.mrl_prb_root_stats(h.value, stats.data_ptr(), 0)))]:
    for _ in range(3): fn()
    print(f"{name:42s} flushed {flushed(fn):6.2f} us | single, warm {noflush_single(fn):6.2f} us | back to back {b2b(fn):6.2f} us", flush=True)
