"""what bounds the time-major DiscountedReturn strips: with / without done flags, affine (two float inputs)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
T, B = 2048, 4096
NS = 5
g = torch.Generator(device="cuda"); g.manual_seed(1)
r = [torch.randn((T, B), device="cuda", generator=g) for _ in range(NS)]
bb = [torch.rand((T, B), device="cuda", generator=g) for _ in range(NS)]
d = [(torch.rand((T, B), device="cuda", generator=g) < 0.005).to(torch.uint8) for _ in range(NS)]
o = [torch.empty((T, B), device="cuda") for _ in range(NS)]
o2 = [torch.empty((T, B), device="cuda") for _ in range(NS)]
lv = torch.randn(B, device="cuda", generator=g)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); fsrc = torch.zeros(64 << 20, device="cuda")
def timeit(fn, n=30):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    tot = 0.0
    for i in range(n):
        flush.fill_(i & 255); fsrc.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(i); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return round(tot / n * 1e3, 2), round(e0.elapsed_time(e1) / n * 1e3, 2)
P = lambda t: t.data_ptr()
n = T * B
for name, fn, byts in [
    ("DR with done (9 B/elem)", lambda i: _ffi.check(L.mrl_discounted_return(P(r[i % NS]), P(d[i % NS]), P(lv), P(o[i % NS]), T, B, 1, 0.99, 0, 0, 0)), 9 * n),
    ("DR no done (8 B/elem)", lambda i: _ffi.check(L.mrl_discounted_return(P(r[i % NS]), None, P(lv), P(o[i % NS]), T, B, 1, 0.99, 0, 0, 0)), 8 * n),
    ("affine a,b (12 B/elem)", lambda i: _ffi.check(L.mrl_affine_scan(P(r[i % NS]), P(bb[i % NS]), P(lv), P(o[i % NS]), T, B, 0, 0, 0)), 12 * n),
    ("GAE done (17 B/elem)", lambda i: _ffi.check(L.mrl_gae(P(r[i % NS]), P(bb[i % NS]), None, P(lv), P(d[i % NS]), P(o[i % NS]), P(o2[i % NS]), T, B, 0.99, 0.95, 0, 0, 0)), 17 * n),
    ("GAE no done (16 B/elem)", lambda i: _ffi.check(L.mrl_gae(P(r[i % NS]), P(bb[i % NS]), None, P(lv), None, P(o[i % NS]), P(o2[i % NS]), T, B, 0.99, 0.95, 0, 0, 0)), 16 * n),
    ("GAE done no ret (13 B/elem)", lambda i: _ffi.check(L.mrl_gae(P(r[i % NS]), P(bb[i % NS]), None, P(lv), P(d[i % NS]), P(o[i % NS]), None, T, B, 0.99, 0.95, 0, 0, 0)), 13 * n),
]:
    for cols, rpu in ((32, 16), (32, 32)):
        _ffi.set_option("strip_cols", cols); _ffi.set_option("strip_rpu", rpu)
        try:
            f, b2b = timeit(fn)
        except Exception as e:
            print(name, cols, rpu, "n/a", e); continue
        print(f"{name:28s} strips {cols}x{rpu}: flushed {f:6.2f} us ({byts / f / 1e3:5.0f} GB/s)  back to back {b2b:6.2f} us ({byts / b2b / 1e3:5.0f} GB/s)", flush=True)
