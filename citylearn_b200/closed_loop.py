"""Closed loops on the device: a policy and the env step each other without the host in between.

The reference's training loops (`citylearn/agents/base.py:155-176`, `citylearn/agents/sac.py:167-196`) are
`actions = agent.predict(observations); observations, reward, ... = env.step(actions)` on the host.  Here the same loop is
captured ONCE as a CUDA graph - `steps_per_replay` iterations of [policy kernels -> `cl_advance_device`] - and replayed: no host
round trip, no launch latency per step.  `cl_advance_device` keeps the time step in a device counter, so a captured launch
advances the episode on every replay (include/citylearn_b200.h).

The policy is the caller's: any `torch.nn.Module` / callable mapping the env's observation tensor `[E, L]` to actions `[E, A]`
with capturable CUDA ops.  `PerBuildingMLP` is a stand-in for the reference's per-building SAC actor (random-init 2 x 256 MLP, tanh
head) used by `bench.py` for BASELINE configs[4]; it is NOT part of the accelerated hot path (PyTorch / cuBLAS, "plumbing").
"""
from __future__ import annotations

from typing import Callable

import torch


class PerBuildingMLP(torch.nn.Module):
    """One actor per building, evaluated for all buildings at once with batched matmuls: obs `[E, B * O]` (every building with the
    same number O of observations) -> actions `[E, B * A_b]` in (-1, 1)."""

    def __init__(self, n_buildings: int, obs_per_building: int, act_per_building: int = 1, hidden: int = 256,
                 dtype: torch.dtype = torch.bfloat16, device='cuda', seed: int = 0):
        super().__init__()
        g = torch.Generator(device='cpu').manual_seed(seed)
        B, O, A, H = n_buildings, obs_per_building, act_per_building, hidden
        self.B, self.O, self.A = B, O, A

        def init(*shape, fan_in):
            return torch.nn.Parameter(((torch.rand(shape, generator=g) * 2 - 1) / fan_in ** 0.5).to(device=device, dtype=dtype))
        self.Op = (O + 7) // 8 * 8                 # layer-1 K padded to 16-byte rows (tensor-core eligible); the padding weights are zero
        w1 = torch.zeros((B, self.Op, H), dtype=dtype, device=device)
        w1[:, :O] = init(B, O, H, fan_in=O).data
        self.w1, self.b1 = torch.nn.Parameter(w1), init(B, 1, H, fan_in=O)
        self.w2, self.b2 = init(B, H, H, fan_in=H), init(B, 1, H, fan_in=H)
        self.w3, self.b3 = init(B, H, A, fan_in=H), init(B, 1, A, fan_in=H)
        self.dtype = dtype

    def forward(self, obs: torch.Tensor) -> torch.Tensor:
        E = obs.shape[0]
        x = torch.nn.functional.pad(obs.view(E, self.B, self.O), (0, self.Op - self.O)).transpose(0, 1).to(self.dtype).contiguous()   # [B, E, Op]
        x = torch.relu(torch.baddbmm(self.b1, x, self.w1))
        x = torch.relu(torch.baddbmm(self.b2, x, self.w2))
        a = torch.tanh(torch.baddbmm(self.b3, x, self.w3))                              # [B, E, A]
        return a.transpose(0, 1).reshape(E, self.B * self.A).float()

    def parameter_count(self) -> int:
        return sum(p.numel() for p in self.parameters())


class ClosedLoop:
    """`steps_per_replay` iterations of [actions = policy(obs); env step] captured as one CUDA graph.

    `run(n_steps)` advances the env by exactly `n_steps` time steps (graph replays, then single captured steps for the remainder)
    and returns the sum of rewards `[E, R]` accumulated on the device.  The env's `time_step` is kept in sync; observations /
    rewards of the last step are in `env.observations` / `loop.reward`."""

    def __init__(self, env, policy: Callable[[torch.Tensor], torch.Tensor], steps_per_replay: int = 8, warmup: int = 3):
        if env._reward_id < 0:
            raise NotImplementedError('ClosedLoop needs a built-in (fused) reward function')
        self.env, self.policy, self.n = env, policy, int(steps_per_replay)
        dev = env.device
        E, A = env.num_envs, max(env.spec.action_dim, 1)
        self.obs = env._obs                       # the env's own observation slab: the kernel overwrites it in place every step
        self.reward = env._reward
        self.act = torch.zeros((E, A), dtype=torch.float32, device=dev)
        self.ret = torch.zeros_like(self.reward)  # running sum of rewards
        self._stream = torch.cuda.Stream(device=dev)
        with torch.cuda.device(dev):
            _ = env.observations                  # materialise the current observation slab (shared-row host paths skip it)
            env._h.device_time_enable(torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.current_stream(dev).synchronize()
            with torch.cuda.stream(self._stream):
                for _ in range(max(warmup, 1)):   # library workspaces (cuBLAS) must exist before the capture
                    with torch.no_grad():
                        self.act.copy_(policy(self.obs))
            self._stream.synchronize()
            self.graph = self._capture(self.n)
            self.graph1 = self._capture(1) if self.n > 1 else self.graph
        self.replays = 0

    def _one(self):
        env = self.env
        with torch.no_grad():
            self.act.copy_(self.policy(self.obs))
        env._h.advance_device(1, self.act.data_ptr(), self.obs.data_ptr(), self.reward.data_ptr(), env._district.data_ptr(),
                              torch.cuda.current_stream(env.device).cuda_stream)
        self.ret.add_(self.reward)

    def _capture(self, n: int) -> torch.cuda.CUDAGraph:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=self._stream):
            for _ in range(n):
                self._one()
        return g

    def run(self, n_steps: int) -> torch.Tensor:
        env = self.env
        left = env.time_steps - 1 - env.time_step
        if n_steps > left:
            raise RuntimeError(f'ClosedLoop.run({n_steps}): only {left} steps left in the episode; call env.reset()')
        full, rest = divmod(int(n_steps), self.n)
        with torch.cuda.device(env.device):
            for _ in range(full):
                self.graph.replay()
            for _ in range(rest):
                self.graph1.replay()
        self.replays += full + rest
        env.time_step += int(n_steps)
        env._obs_current = True
        env._hist_valid = False
        env._kpi_valid = env._kpi_valid and env._kpi_fused
        return self.ret
