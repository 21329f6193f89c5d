"""sweep of the TMA-ring gather (stages x CTAs/SM x chunk) on C3-shaped rows"""
import ctypes as C, os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
L = _ffi.lib()
cap, batch = 200_000, 512
rows = [28224, 28224]
rb = np.asarray(rows, dtype=np.int64)
h = C.c_int64(-1)
_ffi.check(L.mrl_urb_create(cap, len(rows), rb.ctypes.data, None, 1, C.byref(h)))
cols = (C.c_void_p * len(rows))()
_ffi.check(L.mrl_urb_columns(h.value, cols))
for c, r in enumerate(rows):
    _ffi.check(L.mrl_fill_rows(cols[c], cap, r, 1000 + c, 0))
NS = 8
slots = [torch.randint(0, cap, (batch,), device="cuda", dtype=torch.int64) for _ in range(NS)]
outs = [[torch.empty((batch, r), dtype=torch.uint8, device="cuda") for r in rows] for _ in range(NS)]
tabs = [_ffi.ptr_table([o.data_ptr() for o in oo]) for oo in outs]
byts = 2 * batch * sum(rows)
def run(i):
    _ffi.check(L.mrl_urb_gather(h.value, slots[i % NS].data_ptr(), batch, tabs[i % NS], 0))
def timeit():
    for i in range(NS): run(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(40): run(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / 40)
    return best
res = {}
for stages in (2, 4, 8, 12):
    for per_sm in (1, 2, 3, 4, 6, 8):
        for chunk in (4096, 8192, 14336, 28672):
            if stages * ((28224 // -(-28224 // chunk) + 127) // 128 * 128) * per_sm > 220 * 1024: continue
            _ffi.set_option("bulk_stages", stages); _ffi.set_option("bulk_ctas_per_sm", per_sm); _ffi.set_option("bulk_chunk", chunk)
            t = timeit()
            res[f"s{stages}_c{per_sm}_k{chunk}"] = round(t, 2)
best = sorted(res.items(), key=lambda kv: kv[1])
for k, v in best[:12]: print(k, v, "us", round(byts / v / 1e3, 0), "GB/s")
print("default s4_c4_k8192", res.get("s4_c4_k8192"))
# correctness of the best configuration
k = best[0][0]
s_, c_, k_ = int(k.split("_")[0][1:]), int(k.split("_")[1][1:]), int(k.split("_")[2][1:])
_ffi.set_option("bulk_stages", s_); _ffi.set_option("bulk_ctas_per_sm", c_); _ffi.set_option("bulk_chunk", k_)
run(0); torch.cuda.synchronize()
src = torch.empty((cap, rows[0]), dtype=torch.uint8, device="cuda")
L.mrl_memcpy_d2d(src.data_ptr(), cols[0], cap * rows[0], 0); torch.cuda.synchronize()
print("exact", bool(torch.equal(outs[0][0], src[slots[0]])))
json.dump(res, open("gpurun_out/gather_sweep.json", "w"), indent=1)
