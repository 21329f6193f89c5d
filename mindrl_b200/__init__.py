"""mindrl_b200: B200-native (sm_100a) experience data path of MindSpore Reinforcement.

Replay-buffer ops and reverse scans only; see DESIGN.md.  Python is a thin ctypes layer over
include/mindrl_b200.h; there is no CPU fallback.
"""
from . import _ffi
from .core import PriorityReplayBuffer, ReservoirReplayBuffer, ShardedPriorityReplayBuffer, UniformReplayBuffer, get_replay_buffer_elements
from .utils import (DiscountedReturn, NStepBuffer, TensorArray, affine_scan, discounted_return_normalized,
                    discounted_reward, gae, get_vs_and_advantages, normalize_advantage, td_lambda_targets)

__version__ = "0.1.0"
__all__ = ["UniformReplayBuffer", "PriorityReplayBuffer", "ShardedPriorityReplayBuffer", "ReservoirReplayBuffer", "get_replay_buffer_elements",
           "DiscountedReturn", "discounted_reward", "gae", "affine_scan", "get_vs_and_advantages", "NStepBuffer", "normalize_advantage", "TensorArray",
           "td_lambda_targets", "discounted_return_normalized"]
