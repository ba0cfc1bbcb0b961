"""Batch sharding over the GPUs of one box: envs are independent, so rank k simply owns a
contiguous range of global env indices; the only collective is a sum of the episode-statistics vector
(NCCL over NVLink on a GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from ._cabi import STAT_NAMES


def shard_range(global_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """(env_offset, num_envs) of `rank`: contiguous, sizes differ by at most one, union = [0, global_envs)."""
    if not (0 <= rank < world_size) or global_envs < 0:
        raise ValueError("bad rank / world_size / global_envs")
    base, extra = divmod(global_envs, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def allreduce_stats(stats_vec: torch.Tensor, group=None) -> Dict[str, int]:
    """Sum the int64 statistics vector over the ranks of `group` (default group if None) and name the entries.
    Works on any backend: the vector stays on the device it lives on (cuda -> NCCL, cpu -> gloo)."""
    import torch.distributed as dist

    v = stats_vec.clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
    h = v.cpu().tolist()
    return {k: int(h[i]) for i, k in enumerate(STAT_NAMES)}
