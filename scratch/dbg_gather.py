import ctypes as C, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mindrl_b200 import _ffi
from oracle import oracle as O
L = _ffi.lib()
def run(path, row_bytes, cap=2000, n=257, second=8):
    _ffi.set_option("gather_path", path)
    rb = np.asarray([row_bytes, second], dtype=np.int64)
    h = C.c_int64(-1)
    _ffi.check(L.mrl_urb_create(cap, 2, rb.ctypes.data, None, 0, C.byref(h)))
    cols = (C.c_void_p * 2)()
    _ffi.check(L.mrl_urb_columns(h.value, cols))
    _ffi.check(L.mrl_fill_rows(cols[0], cap, row_bytes, 77, 0)); torch.cuda.synchronize(); print("  fill0 ok", flush=True)
    _ffi.check(L.mrl_fill_rows(cols[1], cap, second, 78, 0)); torch.cuda.synchronize(); print("  fill1 ok", flush=True)
    slots = np.random.default_rng(row_bytes).integers(0, cap, size=n)
    ds = torch.as_tensor(slots).cuda()
    out0 = torch.empty((n, row_bytes), dtype=torch.uint8, device="cuda")
    out1 = torch.empty((n, second), dtype=torch.uint8, device="cuda")
    tab = _ffi.ptr_table([out0.data_ptr(), out1.data_ptr()])
    _ffi.check(L.mrl_urb_gather(h.value, ds.data_ptr(), n, tab, 0)); torch.cuda.synchronize(); print("  gather ok", flush=True)
    ok0 = np.array_equal(out0.cpu().numpy(), O.fill_rows(cap, row_bytes, 77)[slots])
    ok1 = np.array_equal(out1.cpu().numpy(), O.fill_rows(cap, second, 78)[slots])
    print("  match", ok0, ok1, flush=True)
    _ffi.check(L.mrl_urb_destroy(h.value))
for path in (1, 2):
    for rbytes in (4096, 28224, 1000, 6, 3):
        for second in (8, 1):
            print("path", path, "row_bytes", rbytes, "second", second, flush=True)
            run(path, rbytes, second=second)
