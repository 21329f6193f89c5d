"""DiscountedReturn and the in-learner reverse scans, on the CUDA library (mrl_discounted_return,
mrl_gae, mrl_affine_scan of include/mindrl_b200.h).

Same surface as mindspore_rl/utils/discounted_return.py:25-92:
    DiscountedReturn(gamma, need_bprop=False, dtype=float32)(reward, done, last_state_value)
"""
import numpy as np

from .. import _carrier as K
from .. import _ffi

_LAYOUTS = {"time_major": _ffi.TIME_MAJOR, "batch_major": _ffi.BATCH_MAJOR,
            _ffi.TIME_MAJOR: _ffi.TIME_MAJOR, _ffi.BATCH_MAJOR: _ffi.BATCH_MAJOR}
_MODES = {"auto": _ffi.SCAN_AUTO, "sequential": _ffi.SCAN_SEQUENTIAL, "parallel": _ffi.SCAN_PARALLEL,
          0: 0, 1: 1, 2: 2}


def _prod(shape):
    return int(np.prod(shape, dtype=np.int64))


def _kind(*xs):
    kinds = {K.classify(x) for x in xs if x is not None}
    return "device" if kinds == {"device"} else "host"


def _prep(x, dtype, kind):
    if x is None:
        return None, None
    if kind == "device":
        t = K.as_device(x, dtype)
        return t, t.data_ptr()
    a = K.as_host(x, dtype)
    return a, a.ctypes.data


def _done_u8(done, kind):
    """done flags as uint8.  Any non-zero value counts as "done" (the reference computes (1 - done) in float,
    discounted_return.py:88; its callers only pass booleans / 0-1 masks, which is what this path supports --
    fractional masks belong to affine_scan, whose b[t] is arbitrary)."""
    if done is None:
        return None, None
    if kind == "device":
        t = done
        if K.is_torch(t) and t.dtype not in (K.torch.bool, K.torch.uint8):
            t = t != 0
        if K.is_torch(t) and t.dtype == K.torch.bool:
            t = t.contiguous().view(K.torch.uint8)
        t = K.as_device(t, np.uint8)
        return t, t.data_ptr()
    if K.is_torch(done):
        done = done.detach().cpu().numpy()
    a = np.ascontiguousarray(np.asarray(done) != 0).view(np.uint8)
    return a, a.ctypes.data


def _empty_like(x, kind):
    if kind == "device":
        return K.device_empty(tuple(x.shape), np.float32)
    return np.empty(tuple(x.shape), dtype=np.float32)


def _out_ptr(x, kind):
    return x.data_ptr() if kind == "device" else x.ctypes.data


class DiscountedReturn:
    """last <- reward[t] + (1 - done[t]) * gamma * last, t = T-1..0 (discounted_return.py:84-92).

    reward [T, B, ...] float32, done [T, B] (broadcast over trailing dims of reward) or both [T];
    last_state_value [B, ...].  Returns an array shaped like reward.  Raises ValueError unless
    0 <= gamma <= 1 (:63-66).  `mode`: 'auto' | 'sequential' (bit-identical to the reference loop
    order) | 'parallel'.
    """

    def __init__(self, gamma, need_bprop=False, dtype=np.float32, mode="auto"):
        if gamma > 1.0 or gamma < 0.0:
            raise ValueError(
                f"The discounted factor should be a number in range [0, 1], but got {gamma}.")
        if K.np_dtype(dtype) != np.float32:
            raise TypeError("DiscountedReturn computes in float32")
        self.gamma = float(gamma)
        self.need_bprop = need_bprop
        self.mode = _MODES[mode]

    def construct(self, reward, done, last_state_value):
        kind = _kind(reward, done, last_state_value)
        r, r_ptr = _prep(reward, np.float32, kind)
        shape = tuple(r.shape)
        if len(shape) == 0:
            raise ValueError("reward needs a time axis")
        T = shape[0]
        dshape = tuple(done.shape) if done is not None else shape
        while len(dshape) > 1 and dshape[-1] == 1:  # [T, B, 1] flags broadcast like [T, B]
            dshape = dshape[:-1]
        if len(dshape) > len(shape) or shape[:len(dshape)] != dshape:
            raise ValueError(f"done {tuple(done.shape)} does not broadcast against reward {shape}")
        B = _prod(dshape[1:])
        E = _prod(shape[1:]) // max(B, 1)
        d, d_ptr = _done_u8(done, kind)
        lv, lv_ptr = None, None
        if last_state_value is not None:
            tail = shape[1:]
            if kind == "device":
                lv = K.as_device(last_state_value, np.float32)
                if K.is_torch(lv):
                    if lv.numel() == _prod(tail):
                        lv = lv.reshape(tail)
                    lv = lv.broadcast_to(tail).contiguous()
                elif _prod(lv.shape) != _prod(tail):
                    raise ValueError("last_state_value does not match reward[0]")
                lv_ptr = lv.data_ptr()
            else:
                h = K.as_host(last_state_value, np.float32)
                if h.size == _prod(tail):
                    h = h.reshape(tail)
                elif h.size == 1:
                    h = h.reshape(())
                lv = np.ascontiguousarray(np.broadcast_to(h, tail))
                lv_ptr = lv.ctypes.data
        out = _empty_like(r, kind)
        fn = (_ffi.lib().mrl_discounted_return if kind == "device"
              else _ffi.lib().mrl_discounted_return_host)
        _ffi.check(fn(r_ptr, d_ptr, lv_ptr, _out_ptr(out, kind), T, B, E, self.gamma,
                      _ffi.TIME_MAJOR, self.mode, K.current_stream()))
        return out

    __call__ = construct


def discounted_reward(rewards, gamma, last_value=None, done=None, layout="batch_major", mode="auto"):
    """PPO's discounted_reward (ppo.py:320-332): v <- r[:, i] + gamma * v over [env, T]."""
    kind = _kind(rewards, done, last_value)
    r, r_ptr = _prep(rewards, np.float32, kind)
    lay = _LAYOUTS[layout]
    if lay == _ffi.BATCH_MAJOR:
        B, T = int(r.shape[0]), _prod(r.shape[1:])
    else:
        T, B = int(r.shape[0]), _prod(r.shape[1:])
    d, d_ptr = _done_u8(done, kind)
    lv, lv_ptr = _prep(last_value, np.float32, kind)
    out = _empty_like(r, kind)
    fn = (_ffi.lib().mrl_discounted_return if kind == "device"
          else _ffi.lib().mrl_discounted_return_host)
    _ffi.check(fn(r_ptr, d_ptr, lv_ptr, _out_ptr(out, kind), T, B, 1, float(gamma), lay,
                  _MODES[mode], K.current_stream()))
    return out


def gae(reward, value, gamma, lam, next_value=None, last_value=None, done=None,
        layout="time_major", mode="auto", want_returns=True, normalize=False, epsilon=1e-8):
    """Generalised advantage estimation.

      m = 1 - done;  delta = r + gamma * V' * m - V;  adv = delta + gamma*lam * m * adv';  ret = adv + V
    PPO form (ppo.py:334-350): next_value = critic(next_state), done=None, layout='batch_major'.
    MAPPO form (mappo.py:478-497): value[t+1] with last_value as bootstrap, done = 1 - mask,
    layout='time_major'.  Returns (advantage, returns or None).
    normalize=True (device arrays): the advantages come back as (adv - mean) / (sqrt(var) + epsilon), PPO's
    _normalized_advantage (ppo.py:361-368), with the moments accumulated by the scan kernel itself."""
    kind = _kind(reward, value, next_value, last_value, done)
    r, r_ptr = _prep(reward, np.float32, kind)
    v, v_ptr = _prep(value, np.float32, kind)
    nv, nv_ptr = _prep(next_value, np.float32, kind)
    lv, lv_ptr = _prep(last_value, np.float32, kind)
    d, d_ptr = _done_u8(done, kind)
    lay = _LAYOUTS[layout]
    if lay == _ffi.BATCH_MAJOR:
        B, T = int(r.shape[0]), _prod(r.shape[1:])
    else:
        T, B = int(r.shape[0]), _prod(r.shape[1:])
    if _prod(v.shape) != T * B:
        raise ValueError("value must have the shape of reward (pass V[T] as last_value)")
    adv = _empty_like(r, kind)
    ret = _empty_like(r, kind) if want_returns else None
    if normalize:
        if kind != "device":
            raise TypeError("gae(normalize=True) takes device arrays")
        _ffi.check(_ffi.lib().mrl_gae_normalized(
            r_ptr, v_ptr, nv_ptr, lv_ptr, d_ptr, adv.data_ptr(), None if ret is None else ret.data_ptr(), T, B,
            float(gamma), float(lam), float(epsilon), lay, _MODES[mode], K.current_stream()))
        return adv, ret
    fn = _ffi.lib().mrl_gae if kind == "device" else _ffi.lib().mrl_gae_host
    _ffi.check(fn(r_ptr, v_ptr, nv_ptr, lv_ptr, d_ptr, _out_ptr(adv, kind),
                  None if ret is None else _out_ptr(ret, kind), T, B, float(gamma), float(lam), lay,
                  _MODES[mode], K.current_stream()))
    return adv, ret


def affine_scan(a, b, x_last=None, layout="time_major", mode="auto"):
    """x[t] = a[t] + b[t] * x[t+1] (device arrays only): V-trace's vs - V
    (vtrace.py:88-93, a = delta_t, b = discount_t * c_t * mask_t) and TD(lambda) (coma.py:489-501)."""
    if _kind(a, b, x_last) != "device":
        raise TypeError("affine_scan takes device arrays")
    at, a_ptr = _prep(a, np.float32, "device")
    bt, b_ptr = _prep(b, np.float32, "device")
    xl, xl_ptr = _prep(x_last, np.float32, "device")
    lay = _LAYOUTS[layout]
    if lay == _ffi.BATCH_MAJOR:
        B, T = int(at.shape[0]), _prod(at.shape[1:])
    else:
        T, B = int(at.shape[0]), _prod(at.shape[1:])
    out = _empty_like(at, "device")
    _ffi.check(_ffi.lib().mrl_affine_scan(a_ptr, b_ptr, xl_ptr, out.data_ptr(), T, B, lay,
                                          _MODES[mode], K.current_stream()))
    return out


def normalize_advantage(advantage, epsilon=1e-8):
    """(adv - mean) / (sqrt(var) + epsilon) over all elements (PPO's _normalized_advantage,
    ppo.py:361-368), device arrays only; moments accumulated in float64 in a fixed order."""
    if _kind(advantage) != "device":
        raise TypeError("normalize_advantage takes a device array")
    x, x_ptr = _prep(advantage, np.float32, "device")
    out = _empty_like(x, "device")
    _ffi.check(_ffi.lib().mrl_normalize(x_ptr, out.data_ptr(), _prod(x.shape), float(epsilon),
                                        K.current_stream()))
    return out


def discounted_return_normalized(reward, done, last_state_value, gamma, epsilon=1e-8, mode="auto"):
    """A2C's returns + normalisation (a2c.py:189-192) as one operation on device arrays:
    DiscountedReturn(gamma)(reward, done, last) followed by (x - mean) / (sqrt(var) + epsilon)."""
    if gamma > 1.0 or gamma < 0.0:
        raise ValueError(f"The discounted factor should be a number in range [0, 1], but got {gamma}.")
    if _kind(reward, done, last_state_value) != "device":
        raise TypeError("discounted_return_normalized takes device arrays")
    r, r_ptr = _prep(reward, np.float32, "device")
    T, B = int(r.shape[0]), _prod(r.shape[1:])
    d, d_ptr = _done_u8(done, "device")
    lv, lv_ptr = _prep(last_state_value, np.float32, "device")
    out = _empty_like(r, "device")
    _ffi.check(_ffi.lib().mrl_discounted_return_normalized(
        r_ptr, d_ptr, lv_ptr, out.data_ptr(), T, B, 1, float(gamma), float(epsilon), _ffi.TIME_MAJOR,
        _MODES[mode], K.current_stream()))
    return out


def td_lambda_targets(rewards, terminated, mask, target_qs, gamma, td_lambda, max_t=None):
    """COMA's build_td_lambda_targets (mindspore_rl/algorithm/coma/coma.py:489-501), batch-major:

        ret[:, t] = td_lambda * gamma * ret[:, t + 1]
                    + mask[:, t] * (rewards[:, t] + (1 - td_lambda) * gamma * target_qs[:, t + 1] * (1 - terminated[:, t]))

    for t = max_t - 1 .. 0 on ret = zeros_like(target_qs).  target_qs is [B, Tq, A] (A agents, Tq > max_t);
    rewards, terminated, mask are [B, Tr, 1] (broadcast over the agents), [B, Tr] or [B, Tr, A].  max_t defaults
    to Tq - 1.  Device arrays in, device array shaped like target_qs out; the whole loop is one kernel of the
    library (mrl_td_lambda) in the reference's operation order, so results are bit-identical."""
    if _kind(rewards, terminated, mask, target_qs) != "device":
        raise TypeError("td_lambda_targets takes device arrays")
    q, q_ptr = _prep(target_qs, np.float32, "device")
    qshape = tuple(q.shape)
    if len(qshape) < 2:
        raise ValueError("target_qs needs [batch, time, ...]")
    Bn, Tq = int(qshape[0]), int(qshape[1])
    A = _prod(qshape[2:])
    r, r_ptr = _prep(rewards, np.float32, "device")
    te, te_ptr = _prep(terminated, np.float32, "device")
    m, m_ptr = _prep(mask, np.float32, "device")
    Tr = int(r.shape[1])
    RA = _prod(r.shape[2:])
    for x in (te, m):
        if tuple(x.shape[:2]) != (Bn, Tr) or _prod(x.shape[2:]) != RA:
            raise ValueError("rewards, terminated and mask must share one shape")
    if int(r.shape[0]) != Bn or RA not in (1, A):
        raise ValueError("rewards must be [batch, time, 1] or [batch, time, agents]")
    max_t = Tq - 1 if max_t is None else int(max_t)
    ret = _empty_like(q, "device")
    _ffi.check(_ffi.lib().mrl_td_lambda(r_ptr, te_ptr, m_ptr, q_ptr, ret.data_ptr(), Bn, Tq, Tr, A, RA, max_t,
                                        float(td_lambda * gamma), float((1 - td_lambda) * gamma),
                                        K.current_stream()))
    return ret
