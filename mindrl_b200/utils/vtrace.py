"""V-trace targets (IMPALA) with the reverse recurrence on the scan kernels.

Same signature and return values as mindspore_rl/algorithm/impala/vtrace.py:48-108
(get_vs_and_advantages).  The element-wise parts (exp, clipping, deltas) stay tensor ops of the
carrier (torch on the GPU); the sequential loop of the reference (vtrace.py:84-93)

    acc = delta_t + discount_t * c_t * acc * mask_t,   t = T-1 .. 0

is one call of mrl_affine_scan with a = delta, b = discount * c * mask.
"""
from .. import _carrier as K
from .discounted_return import affine_scan


def get_vs_and_advantages(log_rhos, discounts, rewards, values, bootstrap_value,
                          clip_rho_threshold, clip_pg_threshold, clip_cs_threshold, masks,
                          mode="auto"):
    """All inputs time-major [T, B] CUDA tensors (bootstrap_value [B]).  Returns (vs, pg_advantages)."""
    torch = K.torch
    if torch is None or not all(K.is_torch(x) and x.is_cuda for x in
                                (log_rhos, discounts, rewards, values, bootstrap_value, masks)):
        raise TypeError("get_vs_and_advantages takes torch CUDA tensors")
    f32 = torch.float32
    log_rhos, discounts, rewards, values, masks = (x.to(f32) for x in (log_rhos, discounts, rewards, values, masks))
    bootstrap_value = bootstrap_value.to(f32)
    rhos = torch.exp(log_rhos)                                                     # :79
    if clip_rho_threshold is not None:
        rhos = torch.minimum(rhos, torch.tensor(clip_rho_threshold, dtype=f32, device=rhos.device))
    cs = rhos if clip_cs_threshold is None else torch.minimum(
        rhos, torch.tensor(clip_cs_threshold, dtype=f32, device=rhos.device))      # :81
    pg_rhos = rhos if clip_pg_threshold is None else torch.minimum(
        rhos, torch.tensor(clip_pg_threshold, dtype=f32, device=rhos.device))      # :82
    values_1 = torch.cat([values[1:], bootstrap_value.unsqueeze(0)], 0)            # :84
    deltas = rhos * (rewards + discounts * values_1 - values)                      # :86
    b = discounts * cs * masks                                                     # :91 (masks are 0/1)
    vs_minus_v_xs = affine_scan(deltas.contiguous(), b.contiguous(), None, layout="time_major", mode=mode)
    vs = vs_minus_v_xs + values                                                    # :99
    target_v_1 = torch.cat([vs[1:], bootstrap_value.unsqueeze(0)], 0)              # :103
    pg_advantage = pg_rhos * (rewards + discounts * target_v_1 - values)           # :105
    return vs, pg_advantage
