"""C2 step launched from the stream vs replayed from a captured CUDA graph (after a flush / back to back)"""
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
pool = 16
prio = torch.randn((pool, batch), device="cuda").abs_() + 1e-3
uni = torch.rand((pool, batch), device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); fsrc = torch.zeros(64 << 20, device="cuda")
def step(i, st):
    _ffi.check(L.mrl_prb_sample(h.value, batch, 0.4, uni[i % pool].data_ptr(), idx.data_ptr(), w.data_ptr(), tab, st))
    _ffi.check(L.mrl_prb_update(h.value, idx.data_ptr(), prio[i % pool].data_ptr(), batch, st))
for i in range(3): step(i, 0)
torch.cuda.synchronize()
graphs = []
side = torch.cuda.Stream()
for i in range(pool):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        step(i, torch.cuda.current_stream().cuda_stream)
    graphs.append(g)
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
for rep in range(2):
    print("stream launches: flushed", flushed(lambda i: step(i, 0)), "us  back to back", b2b(lambda i: step(i, 0)), "us")
    print("graph replay   : flushed", flushed(lambda i: graphs[i % pool].replay()), "us  back to back", b2b(lambda i: graphs[i % pool].replay()), "us", flush=True)
