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


# ------------------------------------------------------------------------------------------------------------------------------
# Building-sharded districts (SURVEY.md §8e "district all-reduce variant"): rank r holds SOME buildings of every env; the per-env
# district sums (citylearn/citylearn.py:1908-1918) - and with them the district term of MARL (citylearn/reward_function.py:132-143) -
# are completed across the ranks inside the step.
# ------------------------------------------------------------------------------------------------------------------------------
class BuildingShardedEnv:
    """This rank's buildings `[first, first + count)` of every env of a district that is split over `world` GPUs.

    `exchange='p2p'` (default): the fused path - the step kernel pushes its partial district sums into every peer's memory over
    NVLink and completes the sums itself (`cl_exchange_*`, include/citylearn_b200.h): ONE launch per step, no collective call.
    `exchange='nccl'`: the two-phase baseline - the step kernel leaves partial sums, `torch.distributed.all_reduce` (NCCL) adds them,
    district-dependent rewards are evaluated afterwards with tensor ops (RewardFunction / MARL only).

    Observations, rewards and actions cover this rank's buildings only (`[E, L_local]`, `[E, B_local]`, `[E, A_local]`); `district`
    is the full-district `[E, 3]` on every rank.  All ranks must step in lock-step."""

    def __init__(self, schema, num_envs: int, device=None, exchange: str = 'p2p', rank: Optional[int] = None, world: Optional[int] = None,
                 connect: bool = True, **kwargs):
        from . import schema as S
        from .env import CityLearnEnv
        if rank is None or world is None:
            rank, world = rank_and_world()
        if exchange not in ('p2p', 'nccl'):
            raise ValueError("exchange must be 'p2p' or 'nccl'")
        self.rank, self.world, self.exchange = rank, world, exchange
        kwargs.pop('central_agent', None)
        kwargs.pop('buildings', None)
        full = S.load(schema, central_agent=False, **kwargs)
        names = [b.name for b in full.buildings]
        self.first, self.count = shard_range(len(names), rank, world)
        if self.count == 0:
            raise ValueError(f'rank {rank} of {world} would own no building out of {len(names)}')
        self.n_buildings_total = len(names)
        def make():
            return CityLearnEnv(schema, num_envs=num_envs, device=device, central_agent=False,
                                buildings=names[self.first:self.first + self.count], debug_trace=(exchange == 'nccl'), **kwargs)
        self.env = make()
        if world > 1 and exchange == 'p2p':
            # ranks wait for each other INSIDE the step: every block of the launch must be resident (cl_exchange_create refuses more than
            # one block per SM).  The default geometry minimises resident threads and may pick small blocks; take the largest ones instead
            import os
            g = self.env._h.geometry()
            sms = torch.cuda.get_device_properties(self.env.device).multi_processor_count
            if g['blocks'] > sms and 'CL_B200_BLOCK_THREADS' not in os.environ:
                self.env.close()
                os.environ['CL_B200_BLOCK_THREADS'] = '480'
                try:
                    self.env = make()
                finally:
                    del os.environ['CL_B200_BLOCK_THREADS']
        self.env.building_offset, self.env.total_buildings = self.first, len(names)
        self._handle_bytes, self._buffer = (None, None)
        if world > 1 and exchange == 'p2p':
            with torch.cuda.device(self.env.device):
                self._handle_bytes, self._buffer = self.env._h.exchange_create(world, rank)
            if connect:
                self.connect_processes()
        if exchange == 'nccl':
            from . import reward_function as rf
            self._marl = isinstance(self.env.reward_function, rf.MARL)
            if not self._marl and type(self.env.reward_function) is not rf.RewardFunction:
                raise NotImplementedError("exchange='nccl' evaluates RewardFunction / MARL only")

    # ---- wiring ----
    def connect_processes(self, group=None):
        """One process per GPU: all-gather the 64-byte IPC handles of the slot arrays and map the peers' arrays."""
        dev = self.env.device
        mine = torch.tensor(list(self._handle_bytes), dtype=torch.uint8, device=dev)
        outs = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(outs, mine, group=group)
        handles = b''.join(bytes(o.cpu().tolist()) for o in outs)
        with torch.cuda.device(dev):
            self.env._h.exchange_connect(handles)
        dist.barrier(group=group)

    @staticmethod
    def connect_in_process(shards):
        """All ranks live in THIS process on different devices (tests, single-process multi-GPU drivers)."""
        bufs = [s._buffer for s in shards]
        devs = [s.env.device.index for s in shards]
        for s in shards:
            with torch.cuda.device(s.env.device):
                s.env._h.exchange_connect_ptrs(bufs, devs)

    # ---- stepping ----
    def reset(self, **kw):
        return self.env.reset(**kw)

    def step(self, actions: torch.Tensor):
        env = self.env
        if self.exchange == 'p2p' or self.world == 1:
            return env.step(actions)
        # two-phase baseline: partial sums -> NCCL all-reduce -> district-dependent rewards
        from . import schema as S
        with torch.cuda.device(env.device):
            a, _ = env._parse_actions(actions)
            env._h.step(a.data_ptr(), env._obs.data_ptr(), None, env._district.data_ptr(), env._trace.data_ptr(), env._stream())
            dist.all_reduce(env._district)
            net = env._trace[:, :, S.DYN['net_electricity_consumption']]
            if self._marl:
                be = -net.double()
                r = torch.sign(be) * 0.01 * be * be * env._district[:, 0:1].double().clamp_min(0.0)
                env._reward.copy_(r.float())
            else:
                m = net.clamp_min(0.0)
                ex = float(env.reward_function.exponent)
                env._reward.copy_(-m if ex == 1.0 else -m.pow(ex))
            env.time_step += 1
            env._obs_current = True
        return env._obs, env._reward, env.terminated, False, {}

    def rollout(self, actions, obs=None, reward=None, district=None):
        if self.exchange != 'p2p' and self.world > 1:
            raise NotImplementedError("rollout() needs exchange='p2p'")
        return self.env.rollout(actions, obs, reward, district)

    @property
    def district(self) -> torch.Tensor:
        return self.env.district

    def exchange_status(self):
        with torch.cuda.device(self.env.device):
            return self.env._h.exchange_status()

    def close(self):
        self.env.close()


# ------------------------------------------------------------------------------------------------------------------------------
# One process, several devices: `CityLearnEnv(..., devices=[...])`
# ------------------------------------------------------------------------------------------------------------------------------
class DeviceShardedEnv:
    """`num_envs` parallel envs sharded by env index over the CUDA devices of this process: one `CityLearnEnv` per device, no
    collective (envs are independent).  `step` takes the fleet's actions `[E, A]` (any device / host) or a list of per-device
    tensors, launches every shard's kernel without synchronising in between, and returns per-device lists (`obs[i]` lives on
    `devices[i]`) - concatenate with `gather()` when one tensor is wanted."""

    def __init__(self, schema, num_envs: int = 1, devices=None, **kwargs):
        from .env import CityLearnEnv
        devices = [torch.device(d) for d in devices]
        if num_envs < len(devices):
            raise ValueError('fewer envs than devices')
        kwargs.pop('device', None)
        self.devices = devices
        self.num_envs = int(num_envs)
        self.spans = [shard_range(self.num_envs, i, len(devices)) for i in range(len(devices))]
        self.envs = [CityLearnEnv(schema, num_envs=c, device=d, **kwargs) for d, (_, c) in zip(devices, self.spans)]
        for e, (o, _) in zip(self.envs, self.spans):
            e.env_offset, e.total_envs = o, self.num_envs

    def __getattr__(self, name):                       # metadata, spaces, names ...: identical on every shard
        if name.startswith('_') or name in ('envs', 'devices', 'spans'):
            raise AttributeError(name)
        return getattr(self.envs[0], name)

    def _split(self, actions):
        if isinstance(actions, (list, tuple)) and len(actions) == len(self.envs) and all(isinstance(a, torch.Tensor) for a in actions):
            return list(actions)
        a = torch.as_tensor(actions)
        return [a[o:o + c].to(e.device, non_blocking=True) for e, (o, c) in zip(self.envs, self.spans)]

    def reset(self, **kw):
        outs = [e.reset(**kw) for e in self.envs]
        return [o for o, _ in outs], {}

    def step(self, actions):
        outs = [e.step(a) for e, a in zip(self.envs, self._split(actions))]
        return [o[0] for o in outs], [o[1] for o in outs], outs[0][2], False, {}

    def rollout(self, actions, obs=None, reward=None, district=None):
        none = [None] * len(self.envs)
        for e, a, o, r, d in zip(self.envs, actions, obs or none, reward or none, district or none):
            e.rollout(a, o, r, d)
        return obs, reward, self.envs[0].terminated

    @staticmethod
    def gather(parts, device=None) -> torch.Tensor:
        device = parts[0].device if device is None else torch.device(device)
        return torch.cat([p.to(device, non_blocking=True) for p in parts], dim=0)

    @property
    def time_step(self) -> int:
        return self.envs[0].time_step

    @property
    def terminated(self) -> bool:
        return self.envs[0].terminated

    def close(self):
        for e in self.envs:
            e.close()
