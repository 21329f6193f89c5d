"""UniformReplayBuffer: same surface as mindspore_rl/core/uniform_replay_buffer.py:50-172, with
P.BufferAppend / P.BufferGetItem / P.BufferSample replaced by the CUDA library behind
include/mindrl_b200.h (mrl_urb_*)."""
import ctypes as C

import numpy as np

from .. import _carrier as K
from .. import _ffi
from ._columns import ColumnSpec, default_carrier


class _BufferAppend:
    """callable kept for code that reaches for `.buffer_append` directly
    (mindspore_rl/algorithm/mappo/mappo_replaybuffer.py:43)"""

    def __init__(self, owner):
        self._owner = owner

    def __call__(self, buffer, exp, count=None, head=None):
        self._owner.insert(exp)
        return buffer


def get_replay_buffer_elements(columns, transpose=False, shape=None):
    """free-function form of MSRL.get_replay_buffer_elements (core/msrl.py:535-557) over a tuple of device
    columns: identity, or the swap of the two leading axes (mrl_transpose01)."""
    if not transpose:
        return tuple(columns)
    out = []
    for e in columns:
        dims = tuple(int(x) for x in e.shape)
        perm = tuple(range(len(dims))) if shape is None else tuple(int(x) for x in shape)
        if len(perm) != len(dims) or len(dims) < 2 or perm[:2] != (1, 0) or perm[2:] != tuple(range(2, len(dims))):
            raise ValueError(f"get_replay_buffer_elements: permutation {shape} of a {len(dims)}-d column is not "
                             "supported (only the swap of the two leading axes, e.g. (1, 0, 2))")
        if K.is_torch(e) and not e.is_contiguous():
            e = e.contiguous()
        o = K.device_empty((dims[1], dims[0]) + dims[2:], (K.np_dtype(e.dtype) if K.is_torch(e) else e.dtype))
        row_bytes = K.nbytes_of(e) // max(dims[0] * dims[1], 1)
        if dims[0] * dims[1]:
            _ffi.check(_ffi.lib().mrl_transpose01(e.data_ptr(), o.data_ptr(), dims[0], dims[1], row_bytes,
                                                  K.current_stream()))
        out.append(o)
    return tuple(out)


class UniformReplayBuffer:
    """FIFO ring of transition columns on the GPU.

    Args mirror the reference (uniform_replay_buffer.py:75): sample_size, capacity, shapes, types.
    `types` may be numpy / torch / mindspore dtypes or strings.  Extra keyword arguments:
    seed (BufferSample seed, reference: get_seed() or 0, :84-89) and carrier ('auto': results live
    on the device, 'numpy': sample()/get_item() return host arrays).
    """

    def __init__(self, sample_size, capacity, shapes, types, seed=0, carrier="auto"):
        self._spec = ColumnSpec(shapes, types)
        self._capacity = int(capacity)
        self._sample_size = int(sample_size)
        self._carrier = default_carrier(carrier)
        # caller-visible storage, zero-initialised (:26-47, :77)
        self.buffer = tuple(K.device_zeros((self._capacity,) + s, d)
                            for s, d in zip(self._spec.shapes, self._spec.dtypes))
        h = C.c_int64(-1)
        _ffi.check(_ffi.lib().mrl_urb_create(
            self._capacity, self._spec.ncols, self._spec.row_bytes_arr.ctypes.data,
            _ffi.ptr_table([b.data_ptr() for b in self.buffer]), int(seed or 0), C.byref(h)))
        self._handle = h.value
        self.buffer_append = _BufferAppend(self)

    # -- cursor state (reference: Parameters `count`, `head`, :79-80) ------------------------
    def _state(self):
        c, h = C.c_int32(), C.c_int32()
        _ffi.check(_ffi.lib().mrl_urb_state(self._handle, C.byref(c), C.byref(h)))
        return c.value, h.value

    @property
    def count(self):
        return self._state()[0]

    @property
    def head(self):
        return self._state()[1]

    # -- ops ----------------------------------------------------------------------------------
    def insert(self, exp):
        """BufferAppend (:101-115). Accepts one transition or a batch of transitions per column;
        returns the whole buffer."""
        n = self._spec.rows_in(exp)
        kind, arrs, tab = self._spec.prepare_inputs(exp)
        fn = _ffi.lib().mrl_urb_append if kind == "device" else _ffi.lib().mrl_urb_append_host
        _ffi.check(fn(self._handle, n, tab, K.current_stream()))
        return self.buffer

    def get_item(self, index):
        """BufferGetItem (:117-128): element at logical position `index` (0 = oldest)."""
        index = int(index)
        outs, tab = self._spec.outputs(None, self._carrier)
        fn = (_ffi.lib().mrl_urb_get_item if self._carrier == "device"
              else _ffi.lib().mrl_urb_get_item_host)
        try:
            _ffi.check(fn(self._handle, index, tab, K.current_stream()))
        except ValueError as e:
            raise IndexError(str(e)) from None
        return tuple(outs)

    def sample(self):
        """BufferSample (:130-139): sample_size rows drawn uniformly with replacement."""
        outs, tab = self._spec.outputs(self._sample_size, self._carrier)
        fn = (_ffi.lib().mrl_urb_sample if self._carrier == "device"
              else _ffi.lib().mrl_urb_sample_host)
        _ffi.check(fn(self._handle, self._sample_size, None, tab, K.current_stream()))
        return tuple(outs)

    def gather(self, slots):
        """rows at explicit physical slots (device int64 tensor or host array)"""
        if K.classify(slots) == "device":
            s = K.as_device(slots, np.int64)
            n = int(K.nbytes_of(s) // 8)
            outs, tab = self._spec.outputs(n, "device")
            _ffi.check(_ffi.lib().mrl_urb_gather(self._handle, s.data_ptr(), n, tab,
                                                 K.current_stream()))
            return tuple(outs)
        s = K.as_host(slots, np.int64)
        d = K.device_empty(s.shape, np.int64)
        st = K.current_stream()
        _ffi.check(_ffi.lib().mrl_memcpy_h2d(d.data_ptr(), s.ctypes.data, s.nbytes, st))
        outs, tab = self._spec.outputs(s.size, "device")
        _ffi.check(_ffi.lib().mrl_urb_gather(self._handle, d.data_ptr(), s.size, tab, st))
        _ffi.check(_ffi.lib().mrl_stream_sync(st))
        return tuple(outs)

    def reset(self):
        """count <- 0 (:141-150); head and data are left as they are."""
        _ffi.check(_ffi.lib().mrl_urb_reset(self._handle, K.current_stream()))
        return True

    def size(self):
        """count.asnumpy() (:152-160)"""
        return np.asarray(self.count, dtype=np.int32)

    def full(self):
        """count >= capacity as a bool array of shape (1,) (:162-172)"""
        return np.asarray([self.count >= self._capacity])

    def get_replay_buffer_elements(self, transpose=False, shape=None):
        """MSRL.get_replay_buffer_elements (core/msrl.py:535-557): every WHOLE column (all `capacity` rows);
        with transpose=True, `shape` is the permutation -- the reference's callers all pass (1, 0, 2), i.e.
        [T, env, d] -> [env, T, d] (ppo_trainer.py:70); permutations that move other axes are not supported."""
        return get_replay_buffer_elements(self.buffer, transpose, shape)

    # -- checkpointing (not in the reference, which never saves its buffers: SURVEY.md section 5) ---
    def state_dict(self):
        """host copy of the columns and the cursor pair"""
        cols = []
        for b in self.buffer:
            cols.append(b.cpu().numpy() if K.is_torch(b) else b.numpy())
        c, h = self._state()
        return {"columns": cols, "count": c, "head": h}

    def load_state_dict(self, state):
        st = K.current_stream()
        for b, src, dt in zip(self.buffer, state["columns"], self._spec.dtypes):
            a = np.ascontiguousarray(src, dtype=dt)
            if a.nbytes != K.nbytes_of(b):
                raise ValueError("checkpoint column does not match this buffer")
            _ffi.check(_ffi.lib().mrl_memcpy_h2d(b.data_ptr(), a.ctypes.data, a.nbytes, st))
        _ffi.check(_ffi.lib().mrl_stream_sync(st))
        _ffi.check(_ffi.lib().mrl_urb_set_state(self._handle, int(state["count"]), int(state["head"])))

    def destroy(self):
        if getattr(self, "_handle", -1) != -1:
            _ffi.lib().mrl_urb_destroy(self._handle)
            self._handle = -1

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __repr__(self):
        return "UniformReplayBuffer<>"
