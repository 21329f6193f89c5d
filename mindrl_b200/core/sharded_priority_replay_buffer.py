"""Sharded prioritised replay (Ape-X style; BASELINE config 5).

One process per GPU.  Each rank owns a sub-buffer with its own sum/min tree.  A sample() call
all-gathers the 16-byte root statistics {sum, min, size, max_priority} of every rank (NCCL over
NVLink; latency only), then each rank draws its strata from its OWN tree and corrects with
importance weights computed against the GLOBAL sum / min / size.  update_priorities() is
rank-local (per-rank learners consume their own samples): no exchange.  The reference has no
sharded buffer; the only collective it uses next to a buffer is AllGather of whole buffers per
episode (example/ppo/src/distribute_policy_1/fragments.py:124-177).

Global index = rank * local_capacity + local_index.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def split_capacity(total_capacity, world_size):
    """rows owned by each rank (equal shares; the total must divide evenly)"""
    if total_capacity % world_size:
        raise ValueError("total capacity must be a multiple of the world size")
    return total_capacity // world_size


def to_global(local_indices, rank, local_capacity):
    return local_indices + rank * local_capacity


def to_local(global_indices, rank, local_capacity):
    return global_indices - rank * local_capacity


def owner_of(global_indices, local_capacity):
    return global_indices // local_capacity


class ShardedPriorityReplayBuffer:
    """`local` is this rank's buffer: mindrl_b200.core.PriorityReplayBuffer on a GPU box.  It
    must provide root_stats(), sample(beta, uniforms=, gstats=), update_priorities(), insert()."""

    def __init__(self, local, local_capacity, group=None, exchange="auto"):
        """exchange: 'p2p' = peer-memory mailboxes written by the sample kernel itself (ranks on one
        node), 'nccl' = all-gather of the 16-byte root statistics, 'auto' = p2p when the local buffer
        supports it."""
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.local = local
        self.local_capacity = int(local_capacity)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.p2p = False
        if exchange in ("auto", "p2p") and hasattr(local, "shard_init") and self.world <= 32:
            mine = local.shard_init(self.rank, self.world)
            parts = [None] * self.world
            dist.all_gather_object(parts, mine, group=group)
            local.shard_connect(b"".join(parts))
            self.p2p = True
        elif exchange == "p2p":
            raise RuntimeError("peer-memory exchange is not available for this buffer")

    def insert(self, *transition):
        return self.local.insert(*transition)

    def global_stats(self):
        """[world, 4] tensor of every rank's {root_sum, root_min, size, max_priority}"""
        mine = self.local.root_stats()
        if not isinstance(mine, torch.Tensor):
            mine = torch.as_tensor(np.asarray(mine, dtype=np.float32))
        mine = mine.reshape(4).contiguous()
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine, group=self.group)
        return torch.stack(parts, 0).contiguous()

    def sample(self, beta, uniforms=None):
        if self.p2p:
            out = self.local.sample(beta, uniforms=uniforms, p2p=True)
        else:
            out = self.local.sample(beta, uniforms=uniforms, gstats=self.global_stats())
        idx = to_global(out[0], self.rank, self.local_capacity)
        return (idx,) + tuple(out[1:])

    def update_priorities(self, indices, priorities):
        return self.local.update_priorities(to_local(indices, self.rank, self.local_capacity),
                                            priorities)
