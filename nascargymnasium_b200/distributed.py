"""Multi-GPU plumbing: one process per GPU, disjoint env shards, no data-path collective.

Envs never interact, so the stepping path needs no exchange step; the only collectives are the
per-iteration gather of rollouts / episode statistics to the learner rank and the reductions a benchmark needs.
Backend is NCCL on GPUs (NVLink 5 / NVSwitch) and gloo in the CPU tests."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np


def shard_range(num_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Rank r owns envs [lo, hi): contiguous, sizes differ by at most one, union = [0, num_envs)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(num_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_track_ids(num_envs: int, n_tracks: int, rank: int, world_size: int) -> np.ndarray:
    """Global assignment env -> track (env_index mod n_tracks over the sorted track list, BASELINE config 4), returned
    for this rank's shard and sorted so that every CTA of the step kernel serves one track."""
    lo, hi = shard_range(num_envs, rank, world_size)
    return np.sort(np.arange(lo, hi, dtype=np.int64) % n_tracks).astype(np.int32)


def reduce_stats(stats: Dict[str, float], device=None, group=None) -> Dict[str, float]:
    """Sum per-rank counters (car_steps, episodes, laps, return_sum, ...) over all ranks."""
    import torch
    import torch.distributed as dist
    keys = sorted(stats)
    t = torch.tensor([float(stats[k]) for k in keys], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return {k: float(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule for multi-GPU numbers: the slowest rank's device time."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_rollout(local, dst: int = 0, group=None):
    """Gather a per-rank rollout tensor (T, E_local, ...) on the learner rank along the env axis (dim 1).
    Returns the concatenated tensor on `dst`, None elsewhere.  Shards may differ in size by one env."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[1]], dtype=torch.int64, device=local.device), group=group)
    emax = int(max(int(s.item()) for s in sizes))
    pad = local
    if local.shape[1] < emax:
        shape = list(local.shape)
        shape[1] = emax - local.shape[1]
        pad = torch.cat([local, local.new_zeros(shape)], dim=1)
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:, :int(s.item())] for b, s in zip(bufs, sizes)], dim=1)
