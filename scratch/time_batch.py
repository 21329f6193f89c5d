"""C2 step, per-kernel and e2e times with the tree prefetch on / off"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
cap, batch = 1_000_000, int(os.environ.get("BATCH", 256))
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
h_idx = torch.empty(batch, dtype=torch.int64).pin_memory()
h_w = torch.empty(batch, dtype=torch.float32).pin_memory()
h_out = [torch.empty((batch, r), dtype=torch.uint8).pin_memory() for r in rows]
h_tab = _ffi.ptr_table([o.data_ptr() for o in h_out])
h_prio = prio.cpu().pin_memory()
def sample(i):
    _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, None, idx.data_ptr(), w.data_ptr(), tab, 0))
def update(i):
    _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), prio[i % 8].data_ptr(), batch, 0))
def step(i):
    sample(i); update(i)
def flushed(fn, n=100):
    tot = 0.0
    for i in range(n):
        flush.fill_(i & 255); fsrc.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return round(tot / n * 1e3, 2)
def b2b(fn, n=200):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / n * 1e3, 2)
def host(n=300, do_flush=True):
    ts = ta = 0.0
    for i in range(n):
        if do_flush:
            flush.fill_(i & 255); fsrc.sum()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _ffi.check(L.mrl_prb_sample_host(h.value, batch, 0.4, None, h_idx.data_ptr(), h_w.data_ptr(), h_tab, 0))
        t1 = time.perf_counter()
        _ffi.check(L.mrl_prb_update_host(h.value, h_idx.data_ptr(), h_prio[i % 8].data_ptr(), batch, 0))
        _ffi.check(L.mrl_stream_sync(0))
        t2 = time.perf_counter()
        ts += t1 - t0; ta += t2 - t1
    return round(ts / n * 1e6, 2), round(ta / n * 1e6, 2)
for pf in (1, 1):
    pass
    _ffi.set_option("tree_prefetch", 1)
    for i in range(5): step(i)
    torch.cuda.synchronize()
    print("host_zero_copy", pf, "flushed us: step", flushed(step), "sample", flushed(sample), "update", flushed(update),
          "| back-to-back us: step", b2b(step), "sample", b2b(sample), "update", b2b(update),
          "| host (sample_host, update_host+sync) flushed", host(), "warm", host(do_flush=False), flush=True)
# floor of a host round trip on this box: an empty stream sync, and a 4-byte copy D2H
t0 = time.perf_counter()
for i in range(1000): _ffi.check(L.mrl_stream_sync(0))
print("idle mrl_stream_sync us", round((time.perf_counter() - t0) * 1e3, 2))
x = torch.zeros(1, device="cuda"); hx = torch.zeros(1).pin_memory()
t0 = time.perf_counter()
for i in range(1000):
    hx.copy_(x, non_blocking=True); torch.cuda.synchronize()
print("4-byte D2H + sync us", round((time.perf_counter() - t0) * 1e3, 2))
t0 = time.perf_counter()
for i in range(1000):
    x.add_(1); torch.cuda.synchronize()
print("tiny kernel + sync us", round((time.perf_counter() - t0) * 1e3, 2))
