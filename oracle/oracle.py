"""ctypes wrapper of the CPU oracle (oracle/mrl_oracle.c) plus a numpy twin of the scans.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under mindrl_b200/ may import this module.

See the header of mrl_oracle.c for what is restated and what pins it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmrl_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "mrl_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "mrl_arith.h")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, i32, u64, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_float
        L.orc_urb_create.restype = vp
        L.orc_urb_create.argtypes = [i64, i32, vp, u64]
        L.orc_urb_destroy.argtypes = [vp]
        L.orc_urb_column.restype = vp
        L.orc_urb_column.argtypes = [vp, i32]
        L.orc_urb_append.argtypes = [vp, i64, vp]
        L.orc_urb_logical_to_slot.restype = i64
        L.orc_urb_logical_to_slot.argtypes = [vp, i64]
        L.orc_urb_get_item.argtypes = [vp, i64, vp]
        L.orc_urb_gather.argtypes = [vp, vp, i64, vp]
        L.orc_urb_gather_mt.argtypes = [vp, vp, i64, vp, i32]
        L.orc_urb_sample.argtypes = [vp, i64, vp, vp]
        L.orc_urb_reset.argtypes = [vp]
        L.orc_urb_state.argtypes = [vp, vp, vp]
        L.orc_prb_create.restype = vp
        L.orc_prb_create.argtypes = [i64, f32, i32, vp, u64, u64]
        L.orc_prb_destroy.argtypes = [vp]
        L.orc_prb_push.argtypes = [vp, vp]
        L.orc_prb_push_batch.argtypes = [vp, i64, vp]
        L.orc_prb_sample.argtypes = [vp, i64, f32, vp, vp, vp, vp]
        L.orc_prb_sample_sharded.argtypes = [vp, i64, f32, vp, vp, i32, vp, vp, vp]
        L.orc_prb_update.argtypes = [vp, vp, vp, i64]
        L.orc_prb_info.argtypes = [vp, vp, vp, vp]
        L.orc_prb_max_priority.restype = f32
        L.orc_prb_max_priority.argtypes = [vp]
        L.orc_prb_sum_nodes.restype = vp
        L.orc_prb_sum_nodes.argtypes = [vp]
        L.orc_prb_min_nodes.restype = vp
        L.orc_prb_min_nodes.argtypes = [vp]
        L.orc_prb_column.restype = vp
        L.orc_prb_column.argtypes = [vp, i32]
        L.orc_prb_root_stats.argtypes = [vp, vp]
        L.orc_discounted_return.argtypes = [vp, vp, vp, vp, i64, i64, i64, f32, i32]
        L.orc_gae.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, i64, f32, f32, i32]
        L.orc_affine_scan.argtypes = [vp, vp, vp, vp, i64, i64, i32]
        f64 = C.c_double
        L.orc_td_lambda.argtypes = [vp, vp, vp, vp, vp, i64, i64, i64, i64, i64, i64, f64, f64]
        L.orc_vtrace.argtypes = [vp, vp, vp, vp, vp, vp, f32, f32, f32, vp, vp, i64, i64]
        L.orc_prb_bench_loop.restype = f64
        L.orc_prb_bench_loop.argtypes = [vp, i64, i64, f32, vp, i64, vp, vp, vp]
        L.orc_urb_bench_loop.restype = f64
        L.orc_urb_bench_loop.argtypes = [vp, i64, i64, vp, i64, vp, vp, i32]
        L.orc_urb_gather_bench.restype = f64
        L.orc_urb_gather_bench.argtypes = [vp, vp, i64, vp, i32, i64]
        L.orc_detpowf.restype = f32
        L.orc_detpowf.argtypes = [f32, f32]
        L.orc_detpowf_array.argtypes = [vp, vp, vp, i64]
        L.orc_libm_pow_array.argtypes = [vp, vp, vp, i64]
        L.orc_uniform01.restype = f32
        L.orc_uniform01.argtypes = [u64, u64]
        L.orc_mix_seed.restype = u64
        L.orc_mix_seed.argtypes = [u64, u64]
        L.orc_fill_rows.argtypes = [vp, i64, i64, u64]
        L.orc_fill_rows_at.argtypes = [vp, vp, i64, i64, u64]
        L.orc_fill_rows_at.restype = None
        L.orc_now.restype = C.c_double
        _lib = L
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data


def _ptr_table(arrays):
    tab = (C.c_void_p * len(arrays))()
    for i, a in enumerate(arrays):
        tab[i] = a if isinstance(a, int) else a.ctypes.data
    return tab


def _row_bytes(shapes, dtypes):
    return [int(np.prod(s, dtype=np.int64)) * np.dtype(d).itemsize for s, d in zip(shapes, dtypes)]


def _as_cols(cols, dtypes):
    return [np.ascontiguousarray(c, dtype=d) for c, d in zip(cols, dtypes)]


class UniformBufferOracle:
    """uniform_replay_buffer.py:75-172 on the CPU oracle."""

    def __init__(self, sample_size, capacity, shapes, dtypes, seed=0):
        self.sample_size, self.capacity = int(sample_size), int(capacity)
        self.shapes = [tuple(s) for s in shapes]
        self.dtypes = [np.dtype(d) for d in dtypes]
        self.row_bytes = _row_bytes(self.shapes, self.dtypes)
        rb = np.asarray(self.row_bytes, dtype=np.int64)
        self._h = lib().orc_urb_create(self.capacity, len(rb), rb.ctypes.data, seed)
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_urb_destroy(self._h)
            self._h = None

    def columns(self):
        out = []
        for i, (s, d, rb) in enumerate(zip(self.shapes, self.dtypes, self.row_bytes)):
            p = lib().orc_urb_column(self._h, i)
            buf = (C.c_uint8 * (self.capacity * rb)).from_address(p)
            out.append(np.frombuffer(buf, dtype=d).reshape((self.capacity,) + s))
        return out

    def insert(self, exp):
        cols = _as_cols(exp, self.dtypes)
        n0 = cols[0].size * cols[0].itemsize // self.row_bytes[0]
        for c, rb in zip(cols, self.row_bytes):
            assert c.size * c.itemsize == n0 * rb
        rc = lib().orc_urb_append(self._h, n0, _ptr_table(cols))
        assert rc == 0

    def _outs(self, n):
        return [np.empty(((n,) if n is not None else ()) + s, dtype=d)
                for s, d in zip(self.shapes, self.dtypes)]

    def get_item(self, index):
        outs = self._outs(None)
        rc = lib().orc_urb_get_item(self._h, int(index), _ptr_table(outs))
        if rc:
            raise IndexError("index out of range")
        return outs

    def logical_to_slot(self, index):
        return lib().orc_urb_logical_to_slot(self._h, int(index))

    def gather(self, slots, threads=1):
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        outs = self._outs(len(slots))
        lib().orc_urb_gather_mt(self._h, slots.ctypes.data, len(slots), _ptr_table(outs), threads)
        return outs

    def sample(self, n=None):
        n = self.sample_size if n is None else n
        slots = np.empty(n, dtype=np.int64)
        outs = self._outs(n)
        rc = lib().orc_urb_sample(self._h, n, slots.ctypes.data, _ptr_table(outs))
        if rc:
            raise RuntimeError("sample size larger than buffer size")
        return slots, outs

    def reset(self):
        lib().orc_urb_reset(self._h)

    def bench_loop(self, steps, row_pool, n=None, threads=1):
        """seconds for `steps` iterations of {append 1 row, sample n} timed inside the C library (bench.py)"""
        n = self.sample_size if n is None else n
        pool = _as_cols(row_pool, self.dtypes)
        npool = pool[0].size * pool[0].itemsize // self.row_bytes[0]
        slots = np.empty(n, dtype=np.int64)
        outs = self._outs(n)
        return lib().orc_urb_bench_loop(self._h, steps, n, _ptr_table(pool), npool, slots.ctypes.data,
                                        _ptr_table(outs), threads)

    def gather_bench(self, slots, threads, reps):
        """seconds for `reps` gathers of the rows at `slots` over `threads` threads, timed in C (bench.py)"""
        slots = np.ascontiguousarray(slots, dtype=np.int64)
        outs = self._outs(len(slots))
        return lib().orc_urb_gather_bench(self._h, slots.ctypes.data, len(slots), _ptr_table(outs), threads, reps)

    def state(self):
        c, h = C.c_int32(), C.c_int32()
        lib().orc_urb_state(self._h, C.addressof(c), C.addressof(h))
        return c.value, h.value


class ReservoirBufferOracle:
    """reservoir_replay_buffer.py:61-100 restated in numpy (test infrastructure).  Vitter's algorithm R as the
    docstring cites it (:27-31): rows 0..capacity-1 fill the reservoir in order; transition number `total`
    (0-based) after that is kept with probability capacity / (total + 1) and then overwrites a uniformly
    drawn slot.  sample(): min(sample_size, size) distinct rows.  Draw order: one U[0,1) per offered
    transition once full, one integer per kept transition, one permutation per sample.  MindSpore's kernel
    is not in /root/reference: parity unpinned, the distribution is what tests check."""

    def __init__(self, capacity, sample_size, shapes, dtypes, seed0=0, seed1=0):
        self.capacity, self.sample_size = int(capacity), int(sample_size)
        self.shapes = [tuple(s) for s in shapes]
        self.dtypes = [np.dtype(d) for d in dtypes]
        self.cols = [np.zeros((self.capacity,) + s, dtype=d) for s, d in zip(self.shapes, self.dtypes)]
        self.rng = np.random.Generator(np.random.PCG64([int(seed0), int(seed1)]))
        self.total = 0

    def push(self, *transition):
        row = [np.asarray(t, dtype=d).reshape(s) for t, s, d in zip(transition, self.shapes, self.dtypes)]
        if self.total < self.capacity:
            slot = self.total
        elif self.rng.random() < self.capacity / float(self.total + 1):
            slot = int(self.rng.integers(0, self.capacity))
        else:
            slot = None
        self.total += 1
        if slot is not None:
            for c, r in zip(self.cols, row):
                c[slot] = r
        return slot

    def size(self):
        return min(self.total, self.capacity)

    def sample(self):
        size = self.size()
        n = min(self.sample_size, size)
        slots = self.rng.permutation(size)[:n]
        return slots, [c[slots] for c in self.cols]


class PriorityBufferOracle:
    """priority_replay_buffer.py:60-138 on the CPU oracle."""

    def __init__(self, alpha, capacity, sample_size, shapes, dtypes, seed0=0, seed1=0):
        self.alpha, self.capacity, self.sample_size = float(alpha), int(capacity), int(sample_size)
        self.shapes = [tuple(s) for s in shapes]
        self.dtypes = [np.dtype(d) for d in dtypes]
        self.row_bytes = _row_bytes(self.shapes, self.dtypes)
        rb = np.asarray(self.row_bytes, dtype=np.int64)
        self._h = lib().orc_prb_create(self.capacity, self.alpha, len(rb), rb.ctypes.data,
                                       seed0, seed1)
        assert self._h

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_prb_destroy(self._h)
            self._h = None

    def insert(self, *transition):
        cols = _as_cols(transition, self.dtypes)
        lib().orc_prb_push(self._h, _ptr_table(cols))

    def insert_batch(self, cols):
        cols = _as_cols(cols, self.dtypes)
        n = cols[0].size * cols[0].itemsize // self.row_bytes[0]
        lib().orc_prb_push_batch(self._h, n, _ptr_table(cols))

    def sample(self, beta, uniforms=None, n=None, gather=True, gstats=None):
        n = self.sample_size if n is None else n
        idx = np.empty(n, dtype=np.int64)
        w = np.empty(n, dtype=np.float32)
        outs = [np.empty((n,) + s, dtype=d) for s, d in zip(self.shapes, self.dtypes)] if gather else None
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float32)
        tab = _ptr_table(outs) if gather else None
        if gstats is None:
            rc = lib().orc_prb_sample(self._h, n, beta, _ptr(u), idx.ctypes.data, w.ctypes.data, tab)
        else:
            g = np.ascontiguousarray(gstats, dtype=np.float32).reshape(-1, 4)
            rc = lib().orc_prb_sample_sharded(self._h, n, beta, _ptr(u), g.ctypes.data, g.shape[0],
                                              idx.ctypes.data, w.ctypes.data, tab)
        if rc:
            raise RuntimeError("PriorityReplayBuffer.sample: the buffer is empty")
        return (idx, w) + (tuple(outs) if gather else ())

    def bench_loop(self, steps, beta, prio_pool, n=None, gather=True):
        """seconds for `steps` iterations of {sample n, update_priorities with prio_pool[step % len]} timed inside
        the C library (bench.py's CPU legs)"""
        n = self.sample_size if n is None else n
        prio = np.ascontiguousarray(prio_pool, dtype=np.float32).reshape(-1, n)
        idx = np.empty(n, dtype=np.int64)
        w = np.empty(n, dtype=np.float32)
        outs = [np.empty((n,) + s, dtype=d) for s, d in zip(self.shapes, self.dtypes)] if gather else None
        return lib().orc_prb_bench_loop(self._h, steps, n, beta, prio.ctypes.data, prio.shape[0], idx.ctypes.data,
                                        w.ctypes.data, _ptr_table(outs) if gather else None)

    def update_priorities(self, indices, priorities):
        i = np.ascontiguousarray(indices, dtype=np.int64)
        p = np.ascontiguousarray(priorities, dtype=np.float32)
        assert i.size == p.size
        lib().orc_prb_update(self._h, i.ctypes.data, p.ctypes.data, i.size)

    def info(self):
        s, h, c = C.c_int64(), C.c_int64(), C.c_int64()
        lib().orc_prb_info(self._h, C.addressof(s), C.addressof(h), C.addressof(c))
        return s.value, h.value, c.value

    def max_priority(self):
        return lib().orc_prb_max_priority(self._h)

    def tree(self):
        _, _, cap2 = self.info()
        n = 2 * cap2
        s = np.frombuffer((C.c_float * n).from_address(lib().orc_prb_sum_nodes(self._h)), dtype=np.float32)
        m = np.frombuffer((C.c_float * n).from_address(lib().orc_prb_min_nodes(self._h)), dtype=np.float32)
        return s, m

    def root_stats(self):
        out = np.empty(4, dtype=np.float32)
        lib().orc_prb_root_stats(self._h, out.ctypes.data)
        return out

    def columns(self):
        out = []
        for i, (s, d, rb) in enumerate(zip(self.shapes, self.dtypes, self.row_bytes)):
            p = lib().orc_prb_column(self._h, i)
            buf = (C.c_uint8 * (self.capacity * rb)).from_address(p)
            out.append(np.frombuffer(buf, dtype=d).reshape((self.capacity,) + s))
        return out


# ---- scans -------------------------------------------------------------------------------
def _f32(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a) != 0, dtype=np.uint8)


def discounted_return(reward, done, last_value, gamma, layout=0):
    """C oracle of discounted_return.py:84-92; time-major [T,B,E...] or batch-major [B,T]."""
    r = _f32(reward)
    d = _u8(done)
    out = np.empty_like(r)
    if layout == 0:
        T = r.shape[0]
        if d is not None:
            B = int(np.prod(d.shape[1:], dtype=np.int64))
        else:
            B = int(np.prod(r.shape[1:], dtype=np.int64))
        E = int(np.prod(r.shape[1:], dtype=np.int64)) // max(B, 1)
    else:
        B, T = r.shape
        E = 1
    lv = None
    if last_value is not None:
        lv = np.ascontiguousarray(np.broadcast_to(np.asarray(last_value, dtype=np.float32),
                                                  r.shape[1:] if layout == 0 else (B,))).reshape(-1)
    rc = lib().orc_discounted_return(r.ctypes.data, _ptr(d), _ptr(lv), out.ctypes.data, T, B, E,
                                     gamma, layout)
    assert rc == 0
    return out


def gae(reward, value, next_value, last_value, done, gamma, lam, layout=0, want_ret=True):
    r, v, nv, lv, d = _f32(reward), _f32(value), _f32(next_value), _f32(last_value), _u8(done)
    if layout == 0:
        T, B = r.shape[0], int(np.prod(r.shape[1:], dtype=np.int64))
    else:
        B, T = r.shape
    adv = np.empty_like(r)
    ret = np.empty_like(r) if want_ret else None
    rc = lib().orc_gae(r.ctypes.data, v.ctypes.data, _ptr(nv), _ptr(lv), _ptr(d), adv.ctypes.data,
                       _ptr(ret), T, B, gamma, lam, layout)
    assert rc == 0
    return adv, ret


def affine_scan(a, b, x_last, layout=0):
    a, b, xl = _f32(a), _f32(b), _f32(x_last)
    if layout == 0:
        T, B = a.shape[0], int(np.prod(a.shape[1:], dtype=np.int64))
    else:
        B, T = a.shape
    out = np.empty_like(a)
    rc = lib().orc_affine_scan(a.ctypes.data, b.ctypes.data, _ptr(xl), out.ctypes.data, T, B, layout)
    assert rc == 0
    return out


def discounted_return_numpy_loop(reward, done, last_state_value, gamma):
    """numpy twin that mirrors discounted_return.py:84-92 op for op (this loop IS the reference's
    CPU path: one python iteration and ~5 small elementwise ops per time step)."""
    reward = np.asarray(reward, dtype=np.float32)
    done = np.asarray(done)
    gamma_t = np.asarray([gamma], dtype=np.float32)            # :73  Tensor([gamma], dtype)
    last_state_value = np.asarray(last_state_value, dtype=np.float32)
    discounted = np.zeros_like(reward)                           # :84
    step = reward.shape[0] - 1                                   # :85
    while step >= 0:                                             # :86
        nd = (1 - done[step].astype(np.int32)).astype(np.float32)
        if nd.ndim and nd.ndim < reward[step].ndim:
            nd = nd.reshape(nd.shape + (1,) * (reward[step].ndim - nd.ndim))
        last_state_value = reward[step] + nd * gamma_t * last_state_value   # :87-89
        discounted[step] = last_state_value                      # :90
        step -= 1                                                # :91
    return discounted


def detpow(x, y):
    x, y = np.broadcast_arrays(_f32(x), _f32(y))
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    out = np.empty_like(x)
    lib().orc_detpowf_array(x.ctypes.data, y.ctypes.data, out.ctypes.data, x.size)
    return out


def libm_pow(x, y):
    x, y = np.broadcast_arrays(_f32(x), _f32(y))
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    out = np.empty_like(x)
    lib().orc_libm_pow_array(x.ctypes.data, y.ctypes.data, out.ctypes.data, x.size)
    return out


def uniform01(seed, ctr):
    return lib().orc_uniform01(seed, ctr)


def set_shard_weights(exact):
    """0: sharded weights against the global sum / min / size; 1: against the true per-shard sampling probability"""
    lib().orc_set_shard_weights(int(bool(exact)))


def td_lambda_targets(rewards, terminated, mask, target_qs, gamma, td_lambda, max_t=None):
    """C oracle of coma.py:489-501; target_qs [B, Tq, A], the others [B, Tr, 1 or A]"""
    q = _f32(target_qs)
    r, te, m = _f32(rewards), _f32(terminated), _f32(mask)
    B, Tq = q.shape[0], q.shape[1]
    A = int(np.prod(q.shape[2:], dtype=np.int64))
    Tr = r.shape[1]
    RA = int(np.prod(r.shape[2:], dtype=np.int64))
    max_t = Tq - 1 if max_t is None else int(max_t)
    ret = np.empty_like(q)
    rc = lib().orc_td_lambda(r.ctypes.data, te.ctypes.data, m.ctypes.data, q.ctypes.data, ret.ctypes.data, B, Tq, Tr, A,
                             RA, max_t, float(gamma), float(td_lambda))
    assert rc == 0
    return ret


def vtrace(log_rhos, discounts, rewards, values, bootstrap_value, clip_rho, clip_pg, clip_cs, masks):
    """C oracle of vtrace.py:79-105 (time-major); a threshold of None disables that clip"""
    lr, d, r, v, b = (_f32(x) for x in (log_rhos, discounts, rewards, values, bootstrap_value))
    m = _f32(masks)
    T, B = lr.shape[0], int(np.prod(lr.shape[1:], dtype=np.int64))
    vs, pg = np.empty_like(lr), np.empty_like(lr)
    inf = float("inf")
    rc = lib().orc_vtrace(lr.ctypes.data, d.ctypes.data, r.ctypes.data, v.ctypes.data, b.ctypes.data, _ptr(m),
                          inf if clip_rho is None else clip_rho, inf if clip_pg is None else clip_pg,
                          inf if clip_cs is None else clip_cs, vs.ctypes.data, pg.ctypes.data, T, B)
    assert rc == 0
    return vs, pg


def fill_rows(nrows, row_bytes, seed):
    a = np.empty((nrows, row_bytes), dtype=np.uint8)
    lib().orc_fill_rows(a.ctypes.data, nrows, row_bytes, seed)
    return a


def fill_rows_at(rows, row_bytes, seed):
    """rows `rows` of fill_rows(...) without materialising the others"""
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    a = np.empty((len(rows), row_bytes), dtype=np.uint8)
    lib().orc_fill_rows_at(a.ctypes.data, rows.ctypes.data, len(rows), row_bytes, seed)
    return a


def now():
    return lib().orc_now()
