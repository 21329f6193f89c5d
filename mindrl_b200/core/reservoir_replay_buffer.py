"""ReservoirReplayBuffer: same surface as mindspore_rl/core/reservoir_replay_buffer.py:61-100 (push,
sample, destroy), with ReservoirReplayBufferCreate / Push / Sample / Destroy replaced by the CUDA library
behind include/mindrl_b200.h: rows live in the ring columns of a uniform buffer (mrl_urb_create), a push
either appends (buffer not full), overwrites a uniformly drawn slot (mrl_urb_write_at) or is dropped, and
sample gathers distinct slots (mrl_urb_gather).

The keep / replace decisions are host arithmetic: one transition per call, exactly like the reference op
(`push(*transition)`, :69-82), so there is nothing for the GPU to decide.  The algorithm is Vitter's
reservoir sampling as the reference's docstring cites it (:27-31): item number `total` (0-based) is kept with
probability capacity / (total + 1) once the reservoir is full, and then replaces a slot drawn uniformly.
[UPSTREAM-UNVERIFIED: MindSpore's kernel source is not under /root/reference; the RNG stream is not a
parity target, the distribution is -- tests/test_host_logic.py checks it.]"""
import ctypes as C

import numpy as np

from .. import _carrier as K
from .. import _ffi
from ._columns import ColumnSpec, default_carrier


class ReservoirReplayBuffer:
    """Args mirror the reference (:61): capacity, sample_size, shapes, dtypes, seed0=0, seed1=0."""

    def __init__(self, capacity, sample_size, shapes, dtypes, seed0=0, seed1=0, carrier="auto"):
        if seed0 < 0 or seed1 < 0:
            raise ValueError("seed0 and seed1 must be non-negative")
        self._spec = ColumnSpec(shapes, dtypes)
        self._capacity = int(capacity)
        self._sample_size = int(sample_size)
        self._carrier = default_carrier(carrier)
        self._rng = np.random.Generator(np.random.PCG64([int(seed0), int(seed1)]))
        self._total = 0  # transitions offered so far
        h = C.c_int64(-1)
        _ffi.check(_ffi.lib().mrl_urb_create(self._capacity, self._spec.ncols, self._spec.row_bytes_arr.ctypes.data,
                                             None, int(seed0), C.byref(h)))
        self._handle = h.value

    # the keep / replace rule, shared with nothing: the oracle restates it independently
    def _slot_for_next(self):
        """-> slot to write the next transition to, -1 to append, or None to drop it"""
        if self._total < self._capacity:
            return -1
        keep = self._rng.random() < self._capacity / float(self._total + 1)
        return int(self._rng.integers(0, self._capacity)) if keep else None

    def push(self, *transition):
        """ReservoirReplayBufferPush (:69-82): one transition per call; returns the handle."""
        if len(transition) == 1 and isinstance(transition[0], (list, tuple)):
            transition = tuple(transition[0])
        kind, arrs, tab = self._spec.prepare_inputs(transition)
        if self._spec.rows_in(transition) != 1:
            raise ValueError("push takes one transition per call")
        slot = self._slot_for_next()
        self._total += 1
        L = _ffi.lib()
        st = K.current_stream()
        if slot == -1:
            fn = L.mrl_urb_append if kind == "device" else L.mrl_urb_append_host
            _ffi.check(fn(self._handle, 1, tab, st))
        elif slot is not None:
            fn = L.mrl_urb_write_at if kind == "device" else L.mrl_urb_write_at_host
            _ffi.check(fn(self._handle, slot, 1, tab, st))
        return np.asarray([self._handle], dtype=np.int64)

    def size(self):
        return min(self._total, self._capacity)

    def sample(self):
        """ReservoirReplayBufferSample (:84-92): min(sample_size, size) DISTINCT rows ("transitions with
        variable-length tensors")."""
        size = self.size()
        n = min(self._sample_size, size)
        if n == 0:
            raise RuntimeError("sample() on an empty ReservoirReplayBuffer")
        slots = np.ascontiguousarray(self._rng.permutation(size)[:n], dtype=np.int64)
        st = K.current_stream()
        d = K.device_empty(slots.shape, np.int64)
        _ffi.check(_ffi.lib().mrl_memcpy_h2d(d.data_ptr(), slots.ctypes.data, slots.nbytes, st))
        outs, tab = self._spec.outputs(n, "device")
        _ffi.check(_ffi.lib().mrl_urb_gather(self._handle, d.data_ptr(), n, tab, st))
        _ffi.check(_ffi.lib().mrl_stream_sync(st))  # `slots` (host) and `d` may go away now
        self.last_slots = slots
        if self._carrier == "device":
            return tuple(outs)
        return tuple(o.cpu().numpy() if K.is_torch(o) else o.numpy() for o in outs)

    def destroy(self):
        if getattr(self, "_handle", -1) != -1:
            h = self._handle
            _ffi.lib().mrl_urb_destroy(h)
            self._handle = -1
            return np.asarray([h], dtype=np.int64)
        return np.asarray([-1], dtype=np.int64)

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __repr__(self):
        return "ReservoirReplayBuffer<>"
