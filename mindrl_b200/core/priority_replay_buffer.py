"""PriorityReplayBuffer: same surface as mindspore_rl/core/priority_replay_buffer.py:28-138, with
_rl_inner_ops.PriorityReplayBuffer{Create,Push,Sample,Update,Destroy} replaced by the CUDA library
behind include/mindrl_b200.h (mrl_prb_*)."""
import ctypes as C

import numpy as np

from .. import _carrier as K
from .. import _ffi
from ._columns import ColumnSpec, default_carrier


class PriorityReplayBuffer:
    """Proportional prioritised replay (sum/min segment tree) on the GPU.

    Positional arguments are the reference's (:60): alpha, capacity, sample_size, shapes, types,
    seed0=0, seed1=0.  Quirks kept: full() is constantly False (:122-124), reset() raises
    ValueError (:126-128), insert/update_priorities/destroy return the int64 handle of shape (1,).
    """

    def __init__(self, alpha, capacity, sample_size, shapes, types, seed0=0, seed1=0,
                 carrier="auto"):
        self._spec = ColumnSpec(shapes, types)
        self._alpha = float(alpha)
        self._capacity = int(capacity)
        self._sample_size = int(sample_size)
        self._carrier = default_carrier(carrier)
        h = C.c_int64(-1)
        _ffi.check(_ffi.lib().mrl_prb_create(
            self._capacity, self._alpha, self._spec.ncols, self._spec.row_bytes_arr.ctypes.data,
            int(seed0), int(seed1), C.byref(h)))
        self._handle = h.value
        self._handle_arr = np.asarray([self._handle], dtype=np.int64)

    # -- PriorityReplayBufferPush (:78-90) ---------------------------------------------------
    def insert(self, *transition):
        if len(transition) == 1 and isinstance(transition[0], (list, tuple)):
            transition = tuple(transition[0])
        n = self._spec.rows_in(transition)
        kind, arrs, tab = self._spec.prepare_inputs(transition)
        L = _ffi.lib()
        if n == 1:
            fn = L.mrl_prb_push if kind == "device" else L.mrl_prb_push_host
            _ffi.check(fn(self._handle, tab, K.current_stream()))
        else:
            fn = L.mrl_prb_push_batch if kind == "device" else L.mrl_prb_push_batch_host
            _ffi.check(fn(self._handle, n, tab, K.current_stream()))
        return self._handle_arr

    # -- PriorityReplayBufferSample (:92-106) ------------------------------------------------
    def sample(self, beta, uniforms=None, gstats=None, p2p=False):
        """-> (indices int64[B], weights float32[B], *columns).
        uniforms: optional B draws in [0,1) to use instead of the buffer's own generator.
        gstats: optional [world,4] device array of all ranks' root stats (sharded buffer).
        p2p: sharded buffer with connected peer mailboxes (shard_init/shard_connect): the root stats
        are exchanged inside the sample kernel; collective over the ranks."""
        n = self._sample_size
        L = _ffi.lib()
        st = K.current_stream()
        if self._carrier == "device":
            idx = K.device_empty((n,), np.int64)
            w = K.device_empty((n,), np.float32)
            outs, tab = self._spec.outputs(n, "device")
            u_ptr, keep = None, None
            if uniforms is not None:
                if K.classify(uniforms) == "device":
                    keep = K.as_device(uniforms, np.float32)
                else:
                    h = K.as_host(uniforms, np.float32)
                    keep = K.device_empty((n,), np.float32)
                    _ffi.check(L.mrl_memcpy_h2d(keep.data_ptr(), h.ctypes.data, h.nbytes, st))
                    _ffi.check(L.mrl_stream_sync(st))
                u_ptr = keep.data_ptr()
            if p2p:
                _ffi.check(L.mrl_prb_sample_sharded_p2p(self._handle, n, float(beta), u_ptr,
                                                        idx.data_ptr(), w.data_ptr(), tab, st))
            elif gstats is None:
                _ffi.check(L.mrl_prb_sample(self._handle, n, float(beta), u_ptr, idx.data_ptr(),
                                            w.data_ptr(), tab, st))
            else:
                g = K.as_device(gstats, np.float32)
                world = int(K.nbytes_of(g) // 16)
                _ffi.check(L.mrl_prb_sample_sharded(self._handle, n, float(beta), u_ptr,
                                                    g.data_ptr(), world, idx.data_ptr(),
                                                    w.data_ptr(), tab, st))
            return (idx, w) + tuple(outs)
        if gstats is not None:
            raise ValueError("gstats needs the device carrier")
        idx = np.empty((n,), dtype=np.int64)
        w = np.empty((n,), dtype=np.float32)
        outs, tab = self._spec.outputs(n, "host")
        u = None if uniforms is None else K.as_host(uniforms, np.float32)
        _ffi.check(L.mrl_prb_sample_host(self._handle, n, float(beta),
                                         None if u is None else u.ctypes.data, idx.ctypes.data,
                                         w.ctypes.data, tab, st))
        return (idx, w) + tuple(outs)

    # -- PriorityReplayBufferUpdate (:108-120) -----------------------------------------------
    def update_priorities(self, indices, priorities):
        L = _ffi.lib()
        st = K.current_stream()
        if K.classify(indices) == "device" and K.classify(priorities) == "device":
            i = K.as_device(indices, np.int64)
            p = K.as_device(priorities, np.float32)
            n = int(K.nbytes_of(i) // 8)
            if K.nbytes_of(p) // 4 != n:
                raise ValueError("indices and priorities disagree in length")
            _ffi.check(L.mrl_prb_update(self._handle, i.data_ptr(), p.data_ptr(), n, st))
        else:
            i = K.as_host(indices, np.int64).reshape(-1)
            p = K.as_host(priorities, np.float32).reshape(-1)
            if i.size != p.size:
                raise ValueError("indices and priorities disagree in length")
            _ffi.check(L.mrl_prb_update_host(self._handle, i.ctypes.data, p.ctypes.data, i.size, st))
        return self._handle_arr

    def full(self):
        """whether the buffer is full: constant False, as in the reference (:122-124)"""
        return np.bool_(False)

    def reset(self):
        raise ValueError("reset() is not supported by PriorityReplayBuffer.")

    def destroy(self):
        """PriorityReplayBufferDestroy (:130-138)"""
        if getattr(self, "_handle", -1) != -1:
            _ffi.lib().mrl_prb_destroy(self._handle)
            self._handle = -1
        return self._handle_arr

    # -- extras (not in the reference) -------------------------------------------------------
    def info(self):
        s, h, c = C.c_int64(), C.c_int64(), C.c_int64()
        _ffi.check(_ffi.lib().mrl_prb_info(self._handle, C.byref(s), C.byref(h), C.byref(c)))
        return s.value, h.value, c.value

    def size(self):
        return self.info()[0]

    def max_priority(self):
        v = C.c_float()
        _ffi.check(_ffi.lib().mrl_prb_max_priority(self._handle, C.byref(v), K.current_stream()))
        return v.value

    def tree(self):
        """(sum_nodes, min_nodes) as host arrays of 2*leaf_capacity floats (root at 1)"""
        cap2 = self.info()[2]
        s = np.empty(2 * cap2, dtype=np.float32)
        m = np.empty(2 * cap2, dtype=np.float32)
        _ffi.check(_ffi.lib().mrl_prb_tree_dump(self._handle, s.ctypes.data, m.ctypes.data,
                                                K.current_stream()))
        return s, m

    # -- checkpointing (not in the reference, which never saves its buffers) ----------------------
    def state_dict(self):
        """host copy of the ring columns, the tree and the scalar state"""
        L = _ffi.lib()
        st = K.current_stream()
        ptrs = (C.c_void_p * self._spec.ncols)()
        _ffi.check(L.mrl_prb_columns(self._handle, ptrs))
        cols = []
        for p, shape, dt, rb in zip(ptrs, self._spec.shapes, self._spec.dtypes, self._spec.row_bytes):
            a = np.empty((self._capacity,) + shape, dtype=dt)
            _ffi.check(L.mrl_memcpy_d2h(a.ctypes.data, p, self._capacity * rb, st))
            cols.append(a)
        size, head, ctr, maxp = C.c_int64(), C.c_int64(), C.c_uint64(), C.c_float()
        _ffi.check(L.mrl_prb_get_state(self._handle, C.byref(size), C.byref(head), C.byref(ctr),
                                       C.byref(maxp), st))
        s, m = self.tree()
        return {"columns": cols, "sum": s, "min": m, "size": size.value, "head": head.value,
                "rng_counter": ctr.value, "max_priority": maxp.value}

    def load_state_dict(self, state):
        L = _ffi.lib()
        st = K.current_stream()
        ptrs = (C.c_void_p * self._spec.ncols)()
        _ffi.check(L.mrl_prb_columns(self._handle, ptrs))
        for p, src, dt, rb in zip(ptrs, state["columns"], self._spec.dtypes, self._spec.row_bytes):
            a = np.ascontiguousarray(src, dtype=dt)
            if a.nbytes != self._capacity * rb:
                raise ValueError("checkpoint column does not match this buffer")
            _ffi.check(L.mrl_memcpy_h2d(p, a.ctypes.data, a.nbytes, st))
        s = np.ascontiguousarray(state["sum"], dtype=np.float32)
        m = np.ascontiguousarray(state["min"], dtype=np.float32)
        if s.size != 2 * self.info()[2] or m.size != s.size:
            raise ValueError("checkpoint tree does not match this buffer")
        _ffi.check(L.mrl_prb_set_state(self._handle, int(state["size"]), int(state["head"]),
                                       int(state["rng_counter"]), float(state["max_priority"]),
                                       s.ctypes.data, m.ctypes.data, st))

    def shard_init(self, rank, world, publish="mutation"):
        """allocates this rank's peer mailbox; returns its 64-byte CUDA IPC handle.
        publish: 'mutation' (default) -- update_priorities / insert publish the root statistics to the peers,
        sample reads what landed a step earlier; 'sample' -- the sample kernel exchanges them itself."""
        buf = C.create_string_buffer(64)
        _ffi.check(_ffi.lib().mrl_prb_shard_init(self._handle, int(rank), int(world), buf))
        _ffi.check(_ffi.lib().mrl_prb_shard_set_mode(self._handle, {"sample": 0, "mutation": 1}[publish]))
        return buf.raw

    def shard_connect(self, all_handles):
        """maps every rank's mailbox; all_handles = world*64 bytes in rank order.  Raises MindRLError
        (MRL_ERR_CUDA) when a peer is not reachable through CUDA IPC (another node)."""
        buf = C.create_string_buffer(bytes(all_handles), len(all_handles))
        _ffi.check(_ffi.lib().mrl_prb_shard_connect(self._handle, buf, K.current_stream()))

    def shard_wait(self):
        """stream-ordered wait (a one-warp kernel) until the peers' statistics the next sharded sample needs have
        arrived: lets the caller put the wait for stragglers where it hurts least"""
        _ffi.check(_ffi.lib().mrl_prb_shard_wait(self._handle, K.current_stream()))

    def shard_diag(self, reset=False):
        """exchange counters of the sharded sample: SM cycles warp 0 spent polling its mailbox, calls that
        polled, calls whose first look missed a peer's word, timeouts (synchronises the stream)"""
        v = [C.c_int64() for _ in range(4)]
        _ffi.check(_ffi.lib().mrl_prb_shard_diag(self._handle, *[C.byref(x) for x in v], int(bool(reset)),
                                                 K.current_stream()))
        return {"wait_cycles": v[0].value, "polled_calls": v[1].value, "waited_calls": v[2].value,
                "timeouts": v[3].value}

    def root_stats(self):
        """device array {root_sum, root_min, size, max_priority} for the sharded exchange"""
        out = K.device_empty((4,), np.float32)
        _ffi.check(_ffi.lib().mrl_prb_root_stats(self._handle, out.data_ptr(), K.current_stream()))
        return out

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def __repr__(self):
        return "PriorityReplayBuffer<>"
