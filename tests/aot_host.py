"""A minimal stand-in for the MindSpore side of an "aot" custom operator (ops.Custom(..., "aot")): it
fabricates (nparam, params, ndims, shapes, dtypes, stream, extra) exactly as MindSpore's AOT kernel mod
does for mindspore_rl/utils/mcts/cpu/cpu_mcts_ops.cc:38-53 and calls the operator through ctypes."""
import ctypes as C

import numpy as np

from mindrl_b200 import _ffi

_NAMES = {np.dtype(np.float32): b"float32", np.dtype(np.float64): b"float64", np.dtype(np.int32): b"int32",
          np.dtype(np.int64): b"int64", np.dtype(np.uint8): b"uint8", np.dtype(np.int8): b"int8",
          np.dtype(np.bool_): b"bool", np.dtype(np.float16): b"float16"}


class AotOp:
    """op = AotOp("PriorityReplayBufferSample", handle=h); rc = op([(ptr, shape, dtype), ...], stream)"""

    def __init__(self, name, **attrs):
        self.L = _ffi.lib()
        self.name = name
        self.extra = self.L.mrl_aot_extra_create()
        for k, v in attrs.items():
            key = k.encode()
            if isinstance(v, bool):
                _ffi.check(self.L.mrl_aot_extra_set_bool(self.extra, key, int(v)))
            elif isinstance(v, int):
                _ffi.check(self.L.mrl_aot_extra_set_int(self.extra, key, v))
            elif isinstance(v, float):
                _ffi.check(self.L.mrl_aot_extra_set_float(self.extra, key, v))
            elif isinstance(v, str):
                _ffi.check(self.L.mrl_aot_extra_set_str(self.extra, key, v.encode()))
            else:  # list of lists of ints
                flat = np.asarray([x for row in v for x in row], dtype=np.int64)
                lens = np.asarray([len(row) for row in v], dtype=np.int32)
                _ffi.check(self.L.mrl_aot_extra_set_int2d(self.extra, key, flat.ctypes.data, lens.ctypes.data, len(v)))
        self.init_rc = None

    def _tables(self, params):
        n = len(params)
        ptrs = (C.c_void_p * n)(*[p[0] for p in params])
        ndims = (C.c_int * n)(*[len(p[1]) for p in params])
        shape_arrs = [(C.c_int64 * max(len(p[1]), 1))(*p[1]) for p in params]
        shapes = (C.POINTER(C.c_int64) * n)(*[C.cast(a, C.POINTER(C.c_int64)) for a in shape_arrs])
        dtypes = (C.c_char_p * n)(*[_NAMES[np.dtype(p[2])] for p in params])
        return n, ptrs, ndims, shapes, dtypes, shape_arrs

    def __call__(self, params, stream=0):
        """params: inputs then outputs, each (device pointer, shape tuple, numpy dtype)"""
        n, ptrs, ndims, shapes, dtypes, keep = self._tables(params)
        if self.init_rc is None:
            self.init_rc = getattr(self.L, self.name + "Init")(ndims, shapes, dtypes, self.extra)
            if self.init_rc != 0:
                return self.init_rc
        return getattr(self.L, self.name)(n, ptrs, ndims, shapes, dtypes, stream, self.extra)

    def close(self):
        if self.extra:
            self.L.mrl_aot_extra_destroy(self.extra)
            self.extra = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def t(x):
    """(pointer, shape, dtype) of a torch CUDA tensor"""
    dt = {"torch.float32": np.float32, "torch.int32": np.int32, "torch.int64": np.int64, "torch.uint8": np.uint8,
          "torch.bool": np.bool_, "torch.float64": np.float64, "torch.int8": np.int8, "torch.float16": np.float16}[str(x.dtype)]
    return (x.data_ptr(), tuple(x.shape), dt)
