from .discounted_return import (DiscountedReturn, affine_scan, discounted_return_normalized, discounted_reward, gae,
                                normalize_advantage, td_lambda_targets)
from .n_step_buffer import NStepBuffer
from .tensor_array import TensorArray
from .vtrace import get_vs_and_advantages

__all__ = ["DiscountedReturn", "discounted_reward", "gae", "affine_scan", "get_vs_and_advantages", "NStepBuffer",
           "normalize_advantage", "TensorArray", "td_lambda_targets", "discounted_return_normalized"]
