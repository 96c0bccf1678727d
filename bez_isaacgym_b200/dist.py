"""Multi-GPU plumbing: one process per GPU (``torch.distributed``, NCCL over NVLink), environments sharded in
contiguous blocks, no collective on the per-step data path.  The only exchanges are tiny SUM all-reduces of
additive fp64 statistics (RunningMeanStd moments, advantage moments, KL) and the PPO gradient all-reduce --
all latency-bound (<= 0.5 MB), so they are packed into as few collectives as possible.

Reference: rl_games' HorovodWrapper (``hvd.allreduce`` average of the normaliser state dicts once per epoch,
``DistributedOptimizer`` for gradients; hook at ``bez_isaacgym/utils/rlgames_utils.py:71-84``).  Differences, both
deliberate: statistics are merged EXACTLY (pivoted sums are additive, so the N-GPU result equals the 1-GPU result
on the union batch up to fp64 re-association) instead of averaging per-rank running stats, and advantages can be
normalised with global instead of per-rank moments (``BASELINE.json`` north_star).
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_distributed(group=None) -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def resolve_group(group=None):
    """The learner's ONE convention (``learner.A2CAgent``): ``group=None`` means "every rank" -- the WORLD group -- whenever
    ``torch.distributed`` runs with more than one rank, and "single process" otherwise.  The low-level helpers below keep the
    explicit meaning (``allreduce_sum_(acc, None)`` is local), so the agent resolves its argument ONCE through this function and
    hands the result to gradients, KL average, parameter broadcast, both RunningMeanStd normalisers and the advantage moments."""
    if group is None and is_distributed(None):
        return dist.group.WORLD
    return group


def shard_range(num_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of global env ids owned by ``rank`` (remainder spread over the first ranks)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(int(num_envs), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(acc: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of an additive statistics vector over ``group``.  ``group=None`` means LOCAL statistics
    (no collective) -- pass ``torch.distributed.group.WORLD`` explicitly to merge over all ranks."""
    if group is not None and is_distributed(group):
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def pack(tensors):
    """Flatten several small fp64 statistic vectors into one buffer -> ONE collective; returns (flat, views)."""
    flat = torch.cat([t.reshape(-1) for t in tensors])
    views, off = [], 0
    for t in tensors:
        views.append(flat[off:off + t.numel()].view(t.shape))
        off += t.numel()
    return flat, views


def allreduce_grads_(params, group=None, average: bool = True, bucket: Optional[torch.Tensor] = None):
    """Gradient all-reduce on ONE flat bucket (the BezKick policy has 124 237 parameters = 497 KB: a single
    latency-bound NCCL call).  ``bucket`` may be a preallocated flat fp32 buffer."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not is_distributed(group):
        return
    n = sum(g.numel() for g in grads)
    if bucket is None or bucket.numel() < n:
        bucket = torch.empty(n, dtype=grads[0].dtype, device=grads[0].device)
    flat = bucket[:n]
    torch.cat([g.reshape(-1) for g in grads], out=flat)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(dist.get_world_size(group))
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()


def broadcast_parameters(module: torch.nn.Module, src: int = 0, group=None):
    """Start-of-training parameter / buffer broadcast (rl_games: hvd.broadcast_parameters)."""
    if not is_distributed(group):
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
