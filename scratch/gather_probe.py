"""what bounds the TMA-ring gather: capacity (TLB reach) and batch (launch boundary) sweep"""
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
rows = [28224, 28224]
rb = np.asarray(rows, dtype=np.int64)
def probe(cap, batch, sorted_slots=False):
    h = C.c_int64(-1)
    _ffi.check(L.mrl_urb_create(cap, len(rows), rb.ctypes.data, None, 1, C.byref(h)))
    cols = (C.c_void_p * len(rows))()
    _ffi.check(L.mrl_urb_columns(h.value, cols))
    for c, r in enumerate(rows):
        _ffi.check(L.mrl_fill_rows(cols[c], cap, r, 1000 + c, 0))
    NS = max(2, min(8, (300 << 20) // (batch * sum(rows)) + 1))
    slots = [torch.randint(0, cap, (batch,), device="cuda", dtype=torch.int64) for _ in range(NS)]
    if sorted_slots:
        slots = [s.sort().values for s in slots]
    outs = [[torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows] for _ in range(NS)]
    tabs = [_ffi.ptr_table([o.data_ptr() for o in oo]) for oo in outs]
    def run(i):
        _ffi.check(L.mrl_urb_gather(h.value, slots[i % NS].data_ptr(), batch, tabs[i % NS], 0))
    for i in range(NS): run(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20): run(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 20)
    byts = 2 * batch * sum(rows)
    print(f"cap {cap:8d} rows ({cap * sum(rows) / 2**30:6.1f} GiB) batch {batch:5d} sorted {int(sorted_slots)}: "
          f"{best:8.2f} us  {byts / best / 1e3:7.0f} GB/s", flush=True)
    del outs, slots
    _ffi.check(L.mrl_urb_destroy(h.value))
    torch.cuda.empty_cache()
for cap in (4096, 32768, 200_000, 1_000_000):
    for batch in (512, 2048, 8192):
        if batch * sum(rows) * 2 > 40 << 30: continue
        probe(cap, batch)
probe(1_000_000, 512, True)
probe(1_000_000, 8192, True)
# plain device copy of the same volume for reference
a = torch.empty(512 * 28224 * 2 * 8, dtype=torch.uint8, device="cuda"); b = torch.empty_like(a)
for n in (512 * 28224 * 2, 512 * 28224 * 2 * 8):
    for _ in range(3): b[:n].copy_(a[:n])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): b[:n].copy_(a[:n])
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e3 / 20
    print(f"torch copy_ of {n / 1e6:.1f} MB: {t:.2f} us {2 * n / t / 1e3:.0f} GB/s")
