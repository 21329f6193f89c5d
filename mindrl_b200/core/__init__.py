from .priority_replay_buffer import PriorityReplayBuffer
from .sharded_priority_replay_buffer import ShardedPriorityReplayBuffer
from .uniform_replay_buffer import UniformReplayBuffer

__all__ = ["UniformReplayBuffer", "PriorityReplayBuffer", "ShardedPriorityReplayBuffer"]
