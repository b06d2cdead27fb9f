"""Multi-GPU plumbing: one process per GPU, parallel envs sharded by env index, no collective on the step path.

Environments are fully independent (the reference has no cross-env term anywhere in `CityLearnEnv.step`,
`citylearn/citylearn.py:978-1056`); the buildings of one env couple only through the district sums, which never leave a
thread block.  So the natural partition is by env: rank r owns envs `[offset_r, offset_r + count_r)` with all B buildings,
tables and parameters replicated.  Collectives appear only OUTSIDE the step path, for fleet-level
statistics (`fleet_reward_stats`): one small all-reduce per logging interval over NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int) -> Tuple[int, int]:
    """(offset, count) of the envs owned by `rank`: contiguous, sizes differ by at most one, covers [0, total_envs)."""
    if total_envs < 0 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError('bad shard arguments')
    base, extra = divmod(total_envs, world_size)
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def rank_and_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def make_sharded_env(schema, total_envs: int, device=None, **kwargs):
    """This rank's shard of a `total_envs`-wide fleet.  The returned env has `.env_offset` / `.total_envs` set."""
    from .env import CityLearnEnv
    rank, world = rank_and_world()
    offset, count = shard_range(total_envs, rank, world)
    if count == 0:
        raise ValueError(f'rank {rank} of {world} would own no env out of {total_envs}')
    if device is None and torch.cuda.is_available():
        # one process per GPU: this rank's device is LOCAL_RANK (torchrun), not whatever happens to be current (cuda:0 on every rank
        # unless the caller ran torch.cuda.set_device)
        import os
        local = os.environ.get('LOCAL_RANK')
        index = int(local) if local is not None else (rank % torch.cuda.device_count() if world > 1 else torch.cuda.current_device())
        device = torch.device('cuda', index)
    env = CityLearnEnv(schema, num_envs=count, device=device, **kwargs)
    env.env_offset, env.total_envs = offset, total_envs
    return env


def fleet_reward_stats(reward_sum: torch.Tensor, reward_min: torch.Tensor, reward_max: torch.Tensor, steps: int,
                       group: Optional[dist.ProcessGroup] = None) -> Dict[str, torch.Tensor]:
    """Fleet-wide episode reward statistics per building from per-env accumulators `[E_local, R]`.

    Mirrors `episode_rewards` (`citylearn/citylearn.py:1034-1040`: min / max / sum / mean over an episode's steps), reduced over
    every env of every rank: three all-reduces of R floats each (SUM, MIN, MAX) - latency-bound, off the step path.
    """
    local_sum = reward_sum.double().sum(dim=0)
    local_n = torch.tensor([float(reward_sum.shape[0])], dtype=torch.float64, device=reward_sum.device)
    local_min = reward_min.min(dim=0).values.double()
    local_max = reward_max.max(dim=0).values.double()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.cat([local_sum, local_n])
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(local_min, op=dist.ReduceOp.MIN, group=group)
        dist.all_reduce(local_max, op=dist.ReduceOp.MAX, group=group)
        local_sum, local_n = packed[:-1], packed[-1:]
    n_envs = local_n.item()
    return {'sum_per_env_mean': local_sum / n_envs, 'mean_per_step': local_sum / (n_envs * max(steps, 1)), 'min': local_min, 'max': local_max,
            'n_envs': int(n_envs)}
