"""e2e of mrl_gae_host at C4 for different piece counts and both layouts"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
T, B = 2048, 4096
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
hr, hv, hd = pin((T, B), torch.float32), pin((T + 1, B), torch.float32), pin((T, B), torch.uint8)
hadv, hret = pin((T, B), torch.float32), pin((T, B), torch.float32)
hr.normal_(); hv.normal_(); hd.zero_()
p = lambda x: x.data_ptr()
for layout in (0, 1):
    for K in (1, 2, 4, 8, 11, 16):
        _ffi.set_option("host_scan_chunks", K)
        ts = []
        for i in range(8):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            _ffi.check(L.mrl_gae_host(p(hr), p(hv), None, p(hv[T]), p(hd), p(hadv), p(hret), T, B, 0.99, 0.95, layout, 0, 0))
            ts.append(time.perf_counter() - t0)
        print("layout", layout, "K", K, "ms", round(min(ts[2:]) * 1e3, 3), flush=True)
# raw PCIe: one direction alone and both at once
d1 = torch.empty((T, B), device="cuda"); d2 = torch.empty((T, B), device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t_(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); return time.perf_counter() - t0
for _ in range(2):
    a = t_(lambda: d1.copy_(hr, non_blocking=True)); b = t_(lambda: hadv.copy_(d2, non_blocking=True))
    def both():
        with torch.cuda.stream(s1): d1.copy_(hr, non_blocking=True)
        with torch.cuda.stream(s2): hadv.copy_(d2, non_blocking=True)
    c = t_(both)
print("33.5 MB h2d", round(33.5 / a / 1e3, 1), "GB/s  d2h", round(33.5 / b / 1e3, 1), "GB/s  both at once", round(67 / c / 1e3, 1), "GB/s total")
