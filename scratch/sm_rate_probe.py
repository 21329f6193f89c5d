"""per-SM throughput of the strip kernels against the number of strips (is ~50 GB/s per SM a per-SM limit or the chip's?)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
T = 2048
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda"); fsrc = torch.zeros(64 << 20, device="cuda")
P = lambda t: t.data_ptr()
_ffi.set_option("tm_kernel", 6)
for B in (256, 512, 1024, 2048, 3072, 4096, 4736, 8192):
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    r = torch.randn((T, B), device="cuda", generator=g); v = torch.randn((T + 1, B), device="cuda", generator=g)
    d = (torch.rand((T, B), device="cuda", generator=g) < 0.005).to(torch.uint8)
    o = torch.empty((T, B), device="cuda"); o2 = torch.empty((T, B), device="cuda")
    res = []
    for name, fn, bpe in (("GAE", lambda: _ffi.check(L.mrl_gae(P(r), P(v), None, P(v[T]), P(d), P(o), P(o2), T, B, 0.99, 0.95, 0, 0, 0)), 17),
                          ("DR", lambda: _ffi.check(L.mrl_discounted_return(P(r), P(d), P(v[T]), P(o), T, B, 1, 0.99, 0, 0, 0)), 9)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        tot = 0.0
        n = 20
        for i in range(n):
            flush.fill_(i & 255); fsrc.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        us = tot / n * 1e3
        strips = (B + 31) // 32
        sms = min(strips, 148)
        byts = bpe * T * B
        res.append(f"{name} {us:6.2f} us {byts / us / 1e3:5.0f} GB/s total, {byts / (us - 2.8) / 1e3 / sms:5.1f} GB/s per busy SM (kernel ~ events - 2.8 us)")
    print(f"B {B:5d} strips {strips:4d}: " + " | ".join(res), flush=True)
