"""Sharded prioritised replay (Ape-X style; BASELINE config 5).

One process per GPU.  Each rank owns a sub-buffer with its own sum/min tree.  A sample() call
all-gathers the 16-byte root statistics {sum, min, size, max_priority} of every rank (NCCL over
NVLink; latency only), then each rank draws its strata from its OWN tree and corrects with
importance weights computed against the GLOBAL sum / min / size.  update_priorities() is
rank-local (per-rank learners consume their own samples): no exchange.  The reference has no
sharded buffer; the only collective it uses next to a buffer is AllGather of whole buffers per
episode (example/ppo/src/distribute_policy_1/fragments.py:124-177).

Global index = rank * local_capacity + local_index.

Importance weights.  Every rank draws its whole batch from its OWN tree, so transition i of rank r is drawn with
probability p_i / (world * S_r), not p_i / S_global.  By default the weights are nevertheless
formed against the global sum, minimum and size -- exact when the shards carry equal priority mass, which uniform
routing of new transitions gives in expectation.  `mindrl_b200._ffi.set_option("shard_weights", 1)` (on every rank)
switches to weights against the true probability, normalised by the largest weight over all ranks: unbiased for any
split of the mass (tests/test_oracle.py::test_exact_shard_weights_undo_the_sampling_probability).
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
    must provide root_stats(), sample(beta, uniforms=, gstats=), update_priorities(), insert().

    Semantics of sample #j on every rank: strata are drawn from the rank's OWN tree; the importance
    weights use, for every rank q, the root statistics q's tree had at q's own sample #j.  With the
    peer-memory exchanges this needs every rank to issue the same sequence of sample / insert /
    update_priorities calls (SPMD learners) on one stream; programs whose ranks mutate their shards at
    different rates use exchange='nccl' (statistics gathered at the call)."""

    def __init__(self, local, local_capacity, group=None, exchange="auto"):
        """exchange:
        'p2p'        peer mailboxes, statistics published by the MUTATING kernels (update_priorities, insert):
                     the sample step reads words that landed a step earlier and does not wait for the slowest rank
        'p2p_sample' peer mailboxes, statistics exchanged by the sample kernel itself (round-1 protocol)
        'nccl'       all-gather of the 16-byte root statistics
        'auto'       'p2p' when every rank is on this node and every mailbox could be mapped (checked
                     collectively), else 'nccl'."""
        if dist is None or not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        if exchange not in ("auto", "p2p", "p2p_sample", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.local = local
        self.local_capacity = int(local_capacity)
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.p2p = False
        self.exchange = "nccl"
        if exchange == "nccl":
            return
        able = hasattr(local, "shard_init") and self.world <= 32
        if not able and exchange != "auto":
            raise RuntimeError("peer-memory exchange is not available for this buffer")
        # every rank must take the same branch: agree on (a) capability, (b) one node, (c) mapping succeeded
        import socket
        info = [None] * self.world
        mine = None
        if able:
            mine = local.shard_init(self.rank, self.world,
                                    publish="sample" if exchange == "p2p_sample" else "mutation")
        dist.all_gather_object(info, (socket.gethostname(), mine), group=group)
        one_node = len({h for h, _ in info}) == 1
        all_able = all(m is not None for _, m in info)
        ok = False
        err = None
        if one_node and all_able:
            try:
                local.shard_connect(b"".join(m for _, m in info))
                ok = True
            except Exception as e:  # a peer could not be mapped (no peer access between the GPUs)
                err = e
        flags = [None] * self.world
        dist.all_gather_object(flags, ok, group=group)
        if all(flags):
            self.p2p = True
            self.exchange = "p2p_sample" if exchange == "p2p_sample" else "p2p"
        elif exchange != "auto":
            raise RuntimeError(f"peer-memory exchange could not be set up on every rank "
                               f"(one node: {one_node}, mapped here: {ok}): {err}")
        # (a rank whose mapping succeeded while another's failed simply never uses its mailbox)

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

    def wait_for_peers(self):
        """optional: enqueue the wait for the peers' statistics of the NEXT sample now (e.g. before the learner's
        forward/backward pass), so that the sample itself finds them in place; no-op for 'nccl' / 'p2p_sample'"""
        if self.p2p and self.exchange == "p2p":
            self.local.shard_wait()

    def exchange_diag(self, reset=False):
        """counters of the peer-memory exchange (see PriorityReplayBuffer.shard_diag); None for 'nccl'"""
        return self.local.shard_diag(reset) if self.p2p else None
