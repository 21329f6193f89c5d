"""V-trace targets (IMPALA) on the CUDA library.

Same signature and return values as mindspore_rl/algorithm/impala/vtrace.py:48-108
(get_vs_and_advantages).  Everything -- exp and the three clips (:79-82), the deltas (:84-86), the
reverse recurrence (:88-93), vs (:99) and the policy-gradient advantages (:103-105) -- runs inside
mrl_vtrace (csrc/scan.cu): one kernel for IMPALA-sized unrolls, three passes for long narrow
trajectories.  PyTorch only carries the device memory.
"""
import math

import numpy as np

from .. import _carrier as K
from .. import _ffi
from .discounted_return import _MODES


def get_vs_and_advantages(log_rhos, discounts, rewards, values, bootstrap_value,
                          clip_rho_threshold, clip_pg_threshold, clip_cs_threshold, masks,
                          mode="auto"):
    """All inputs time-major [T, B] device arrays (bootstrap_value [B]); a threshold of None disables
    that clip; masks may be None (= 1).  Returns (vs, pg_advantages)."""
    xs = (log_rhos, discounts, rewards, values, bootstrap_value) + (() if masks is None else (masks,))
    if not all(K.classify(x) == "device" for x in xs):
        raise TypeError("get_vs_and_advantages takes device arrays")
    lr, d, r, v, bv = (K.as_device(x, np.float32) for x in (log_rhos, discounts, rewards, values, bootstrap_value))
    m = None if masks is None else K.as_device(masks, np.float32)
    T = int(lr.shape[0])
    B = int(np.prod(lr.shape[1:], dtype=np.int64))
    for name, x in (("discounts", d), ("rewards", r), ("values", v)) + ((("masks", m),) if m is not None else ()):
        if int(np.prod(x.shape, dtype=np.int64)) != T * B:
            raise ValueError(f"{name} must have the shape of log_rhos")
    if int(np.prod(bv.shape, dtype=np.int64)) != B:
        raise ValueError("bootstrap_value must have the shape of log_rhos[0]")
    vs = K.device_empty(tuple(lr.shape), np.float32)
    pg = K.device_empty(tuple(lr.shape), np.float32)
    inf = math.inf
    _ffi.check(_ffi.lib().mrl_vtrace(
        lr.data_ptr(), d.data_ptr(), r.data_ptr(), v.data_ptr(), bv.data_ptr(), None if m is None else m.data_ptr(),
        inf if clip_rho_threshold is None else float(clip_rho_threshold),
        inf if clip_pg_threshold is None else float(clip_pg_threshold),
        inf if clip_cs_threshold is None else float(clip_cs_threshold),
        vs.data_ptr(), pg.data_ptr(), T, B, _MODES[mode], K.current_stream()))
    return vs, pg
