"""Multi-GPU plumbing: games are independent, so they shard by global env index with no collective on
the data path.  Each rank owns a contiguous range ``[env_base, env_base + num_local)``;
the device-side Philox streams are keyed by the GLOBAL env id, so a game's trajectory does not depend on
which GPU hosts it.  The only collective is the optional end-of-run sum of the statistics counters.

Pure host logic (works with the gloo backend on CPU tensors as well as NCCL on CUDA tensors).
"""
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist

# device counters of sx_step_all (include/stratego_b200.h): "steps" = actions processed (accepted or not), "resets" =
# setups drawn inside the step (re-draws of unplayable setups included)
STAT_NAMES = ("games_finished", "player_1_wins", "player_2_wins", "invalid_endings", "illegal_actions",
              "steps", "attacks", "resets")


def shard_bounds(global_envs: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the global env indices owned by `rank`: contiguous, sizes differ by at most one."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    if global_envs < 0:
        raise ValueError("global_envs must be >= 0")
    base, extra = divmod(global_envs, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


@dataclass(frozen=True)
class Shard:
    rank: int
    world_size: int
    local_rank: int
    env_base: int
    num_local: int
    global_envs: int


def current_shard(global_envs: Optional[int] = None, envs_per_rank: Optional[int] = None) -> Shard:
    """Shard of this process from torch.distributed (if initialised) or the torchrun environment variables.
    Give either the global env count (strong split) or the per-rank count (weak scaling)."""
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    if (global_envs is None) == (envs_per_rank is None):
        raise ValueError("give exactly one of global_envs / envs_per_rank")
    if envs_per_rank is not None:
        return Shard(rank, world, local_rank, rank * envs_per_rank, envs_per_rank, world * envs_per_rank)
    lo, hi = shard_bounds(global_envs, world, rank)
    return Shard(rank, world, local_rank, lo, hi - lo, global_envs)


def reduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """Sums the int64 statistics counters over all ranks (in place); a no-op without a process group."""
    if stats.dtype != torch.int64:
        raise TypeError("stats must be int64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def max_over_ranks(value: float, device=None, group=None) -> float:
    """max of a host scalar over all ranks (used for 'slowest rank' timing)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def stats_dict(stats: torch.Tensor) -> dict:
    vals = stats.detach().cpu().tolist()
    return {n: int(v) for n, v in zip(STAT_NAMES, vals) if not n.startswith("reserved")}
