from .priority_replay_buffer import PriorityReplayBuffer
from .reservoir_replay_buffer import ReservoirReplayBuffer
from .sharded_priority_replay_buffer import ShardedPriorityReplayBuffer
from .uniform_replay_buffer import UniformReplayBuffer, get_replay_buffer_elements

__all__ = ["UniformReplayBuffer", "PriorityReplayBuffer", "ShardedPriorityReplayBuffer", "ReservoirReplayBuffer", "get_replay_buffer_elements"]
