"""Env-batch sharding across GPUs: contiguous global-id ranges, no data-path collective.

Environments are independent, so rank r of R owns the global env ids
[r*E/R, (r+1)*E/R); the device RNG is keyed by the global id, so trajectories do not depend on R.
The only exchange is the all-reduce of the episode-statistics vector (NCCL over NVLink on GPUs; the
same code runs on gloo for the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("env_steps", "agent_steps", "episodes", "episode_len_sum", "crashes", "apples", "unresolved",
             "fear_nonzero", "return_sum", "fear_sum", "fear_tasks")


def shard_range(global_envs: int, rank: int, world: int) -> Tuple[int, int]:
    """(env_id_base, num_envs) of `rank`; ranges tile [0, global_envs) exactly, sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    lo = global_envs * rank // world
    hi = global_envs * (rank + 1) // world
    return lo, hi - lo


def allreduce_stats(stats: Dict[str, float], device: Optional[torch.device] = None, group=None) -> Dict[str, float]:
    """Sum a gw_get_stats() dict over all ranks (one 10-element fp64 all-reduce)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(stats)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    vec = torch.tensor([float(stats[k]) for k in STAT_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    out = {k: float(v) for k, v in zip(STAT_KEYS, vec.tolist())}
    for k in STAT_KEYS[:8] + STAT_KEYS[10:]:
        out[k] = int(round(out[k]))
    return out


def max_over_ranks(value: float, device: Optional[torch.device] = None, group=None) -> float:
    """Timing rule: a multi-GPU time is the max over ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
