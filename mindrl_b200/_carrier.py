"""Tensor carriers: where the bytes of a call live.

PyTorch is optional.  A torch CUDA tensor is passed by data_ptr() (device path, results come back
as torch CUDA tensors on the current torch stream); numpy arrays use the *_host entry points of the
ABI (the library stages host<->device copies) and results come back as numpy arrays; without torch
the device-resident storage is exposed through DeviceArray.
"""
import ctypes as C

import numpy as np

from . import _ffi

try:  # optional carrier
    import torch
except Exception:  # pragma: no cover
    torch = None

_NAMES = {
    "float32": np.float32, "float": np.float32, "float64": np.float64, "double": np.float64,
    "float16": np.float16, "half": np.float16, "int8": np.int8, "uint8": np.uint8,
    "int16": np.int16, "int32": np.int32, "int": np.int32, "int64": np.int64, "long": np.int64,
    "bool": np.bool_, "bool_": np.bool_, "uint16": np.uint16, "uint32": np.uint32,
    "uint64": np.uint64,
}


def np_dtype(t):
    """numpy / torch / mindspore dtype objects and strings -> numpy dtype.
    (the reference passes mindspore.dtype objects, uniform_replay_buffer.py:26-47)"""
    if torch is not None and isinstance(t, torch.dtype):
        return np.dtype(str(t).replace("torch.", "").replace("bool", "bool_"))
    try:
        return np.dtype(t)
    except TypeError:
        pass
    name = str(t).lower().replace("mindspore.", "").replace("torch.", "")
    if name in _NAMES:
        return np.dtype(_NAMES[name])
    raise TypeError(f"unsupported dtype {t!r}")


def torch_dtype(npdt):
    return getattr(torch, "bool" if np.dtype(npdt) == np.bool_ else np.dtype(npdt).name)


def has_torch_cuda():
    return torch is not None and torch.cuda.is_available()


def is_torch(x):
    return torch is not None and isinstance(x, torch.Tensor)


def current_stream():
    if has_torch_cuda():
        return torch.cuda.current_stream().cuda_stream
    return 0


class DeviceArray:
    """minimal device buffer used when torch is not the carrier"""

    def __init__(self, shape, dtype, ptr=None, owner=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self._own = ptr is None
        self._owner = owner
        if ptr is None:
            p = C.c_void_p()
            _ffi.check(_ffi.lib().mrl_malloc(C.byref(p), max(self.nbytes, 1)))
            ptr = p.value
        self.ptr = ptr

    def data_ptr(self):
        return self.ptr

    def numpy(self, stream=0):
        out = np.empty(self.shape, dtype=self.dtype)
        _ffi.check(_ffi.lib().mrl_memcpy_d2h(out.ctypes.data, self.ptr, self.nbytes, stream))
        _ffi.check(_ffi.lib().mrl_stream_sync(stream))
        return out

    def copy_from(self, arr, stream=0):
        arr = np.ascontiguousarray(arr, dtype=self.dtype)
        assert arr.nbytes == self.nbytes
        _ffi.check(_ffi.lib().mrl_memcpy_h2d(self.ptr, arr.ctypes.data, self.nbytes, stream))
        _ffi.check(_ffi.lib().mrl_stream_sync(stream))

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "ptr", None):
            try:
                _ffi.lib().mrl_free(self.ptr)
            except Exception:
                pass
            self.ptr = None


def device_empty(shape, dtype):
    if has_torch_cuda():
        return torch.empty(tuple(shape), dtype=torch_dtype(dtype), device="cuda")
    return DeviceArray(shape, dtype)


def device_zeros(shape, dtype):
    if has_torch_cuda():
        return torch.zeros(tuple(shape), dtype=torch_dtype(dtype), device="cuda")
    a = DeviceArray(shape, dtype)
    _ffi.check(_ffi.lib().mrl_memset(a.ptr, 0, a.nbytes, 0))
    _ffi.check(_ffi.lib().mrl_stream_sync(0))
    return a


def classify(x):
    """'device' for torch CUDA tensors / DeviceArray, 'host' for everything numpy can hold"""
    if isinstance(x, DeviceArray):
        return "device"
    if is_torch(x):
        return "device" if x.is_cuda else "host"
    return "host"


def as_host(x, dtype):
    if is_torch(x):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=dtype)


def as_device(x, dtype):
    """torch CUDA tensor (contiguous, right dtype) or DeviceArray"""
    if isinstance(x, DeviceArray):
        if x.dtype != np.dtype(dtype):
            raise TypeError(f"expected {np.dtype(dtype)}, got {x.dtype}")
        return x
    t = x
    want = torch_dtype(dtype)
    if t.dtype != want:
        t = t.to(want)
    return t.contiguous()


def nbytes_of(x):
    if isinstance(x, DeviceArray):
        return x.nbytes
    if is_torch(x):
        return x.numel() * x.element_size()
    return x.nbytes
