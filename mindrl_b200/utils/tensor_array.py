"""TensorArray: same surface as mindspore_rl/utils/tensor_array.py:25-190 (write, read, stack, size, clear,
close), the per-step trajectory staging of A2C / A3C / PG / IMPALA (a2c.py:116-164: write(t, ...) inside the
rollout loop, stack() before DiscountedReturn).  Storage is one device array [capacity, *element_shape]; a
write is a one-row copy into it (device-to-device, or host-to-device for host values), so that stack() is a
single contiguous copy and the scans can read the staged trajectory time-major as it lies.

Semantics kept from the MindSpore ops the reference wraps (TensorArrayWrite/Read/Stack/Clear/Close):
fixed-size arrays (dynamic_size=False) reject indices >= size and stack() returns all `size` rows (unwritten
rows are zeros); dynamic arrays grow and stack() returns the rows up to the highest index written; clear()
keeps the instance and zeroes it; after close() every call raises.  [UPSTREAM-UNVERIFIED: the kernels are
in MindSpore, not under /root/reference; the docstring examples at :41-58 are run by the tests.]"""
import numpy as np

from .. import _carrier as K
from .. import _ffi


class TensorArray:
    def __init__(self, dtype, element_shape, dynamic_size=True, size=0, name="TA"):
        self._dtype = K.np_dtype(dtype)
        if self._dtype.kind not in "biuf":
            raise TypeError(f"For 'TensorArray', dtype must be a number type or bool, got {dtype!r}")
        if not isinstance(size, (int, np.integer)) or size < 0:
            raise ValueError(f"For 'TensorArray', size must be an int >= 0, got {size!r}")
        self._shape = tuple(int(x) for x in element_shape)
        self._dynamic = bool(dynamic_size)
        self._max = int(size)
        self.name = name
        self._row_bytes = int(np.prod(self._shape, dtype=np.int64)) * self._dtype.itemsize
        cap = self._max if not self._dynamic else max(self._max, 16)
        self._buf = K.device_zeros((cap,) + self._shape, self._dtype)
        self._cap = cap
        self._size = 0
        self._closed = False

    def _check(self):
        if self._closed:
            raise RuntimeError(f"TensorArray {self.name!r} is closed")

    def _ptr(self, index):
        return self._buf.data_ptr() + index * self._row_bytes

    def _grow(self, need):
        cap = self._cap
        while cap < need:
            cap *= 2
        new = K.device_zeros((cap,) + self._shape, self._dtype)
        st = K.current_stream()
        if self._size:
            _ffi.check(_ffi.lib().mrl_memcpy_d2d(new.data_ptr(), self._buf.data_ptr(), self._size * self._row_bytes, st))
            _ffi.check(_ffi.lib().mrl_stream_sync(st))  # the old storage goes away with this call
        self._buf, self._cap = new, cap

    def write(self, index, value):
        """TensorArrayWrite (:86-99): value -> position index; returns True."""
        self._check()
        index = int(index)
        if index < 0 or (not self._dynamic and index >= self._max):
            raise IndexError(f"TensorArray {self.name!r}: index {index} out of range [0, {self._max})")
        if index >= self._cap:
            self._grow(index + 1)
        st = K.current_stream()
        if K.classify(value) == "device":
            v = K.as_device(value, self._dtype)
            if K.nbytes_of(v) != self._row_bytes:
                raise ValueError(f"TensorArray {self.name!r}: value does not have element shape {self._shape}")
            _ffi.check(_ffi.lib().mrl_memcpy_d2d(self._ptr(index), v.data_ptr(), self._row_bytes, st))
        else:
            v = K.as_host(value, self._dtype)
            if v.nbytes != self._row_bytes:
                raise ValueError(f"TensorArray {self.name!r}: value does not have element shape {self._shape}")
            _ffi.check(_ffi.lib().mrl_memcpy_h2d(self._ptr(index), v.ctypes.data, self._row_bytes, st))
            _ffi.check(_ffi.lib().mrl_stream_sync(st))  # `v` may be a temporary
        self._size = max(self._size, index + 1)
        return True

    def read(self, index):
        """TensorArrayRead (:101-112): the value at position index (a copy)."""
        self._check()
        index = int(index)
        if index < 0 or index >= self._size:
            raise IndexError(f"TensorArray {self.name!r}: index {index} out of range [0, {self._size})")
        out = K.device_empty(self._shape, self._dtype)
        _ffi.check(_ffi.lib().mrl_memcpy_d2d(out.data_ptr(), self._ptr(index), self._row_bytes, K.current_stream()))
        return out

    def size(self):
        """TensorArraySize (:160-168)"""
        self._check()
        return self._size

    def stack(self):
        """TensorArrayStack (:148-158): [n, *element_shape]; n = size for a fixed-size array, else the number
        of positions written so far."""
        self._check()
        n = self._max if not self._dynamic else self._size
        out = K.device_empty((n,) + self._shape, self._dtype)
        if n:
            _ffi.check(_ffi.lib().mrl_memcpy_d2d(out.data_ptr(), self._buf.data_ptr(), n * self._row_bytes,
                                                 K.current_stream()))
        return out

    def view(self):
        """zero-copy view of the staged rows (not in the reference): what the scans can read directly"""
        self._check()
        n = self._max if not self._dynamic else self._size
        return self._buf[:n] if K.is_torch(self._buf) else self._buf

    def clear(self):
        """TensorArrayClear (:133-146): keep the instance, drop the data, size <- 0."""
        self._check()
        if self._cap:
            _ffi.check(_ffi.lib().mrl_memset(self._buf.data_ptr(), 0, self._cap * self._row_bytes, K.current_stream()))
        self._size = 0
        return True

    def close(self):
        """TensorArrayClose (:114-131): free everything; the array is unusable afterwards."""
        self._check()
        self._buf = None
        self._closed = True
        return True

    def __repr__(self):
        return f"TensorArray<{self.name}>"
