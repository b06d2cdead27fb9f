"""Environment wrappers (mirror of `citylearn/wrappers.py:15-238, 516-621`) with their arithmetic fused into the CUDA kernels.

The reference wraps the env in Gymnasium `ObservationWrapper` / `ActionWrapper` objects that re-walk every building's
observation dict on the host each step.  Here a wrapper only *configures* the device-side district
(`CityLearnEnv.configure_transforms` -> `cl_set_transforms`): the periodic sin / cos expansion, the min-max scaling and the
clipping happen where the observation row is written (the helper warp builds the row once per block and step), the action
de-normalisation where the action is fetched.  The wrapped env then returns already transformed tensors, for any `num_envs`.

Same class names, constructor argument (`env`) and `observation_space` / `action_space` / `observation_names` properties as the
reference; `env.unwrapped` keeps reporting the raw spaces and names.  Constructing a wrapper starts a fresh episode (the
reference's wrappers are also applied before the first `reset`).
"""
from __future__ import annotations

from typing import List

import numpy as np

from .spaces import Box

__all__ = ['Wrapper', 'ClippedObservationWrapper', 'NormalizedObservationWrapper', 'NormalizedActionWrapper', 'NormalizedSpaceWrapper',
           'StableBaselines3ObservationWrapper', 'StableBaselines3ActionWrapper', 'StableBaselines3RewardWrapper', 'StableBaselines3Wrapper']


class Wrapper:
    """Minimal stand-in for `gymnasium.Wrapper`: delegates everything it does not override to the wrapped env."""

    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name == 'env':
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, actions):
        return self.env.step(actions)


class ClippedObservationWrapper(Wrapper):
    """Observations are clipped to the observation-space limits (`citylearn/wrappers.py:15-37`)."""

    def __init__(self, env):
        super().__init__(env)
        self.unwrapped.configure_transforms(observation_transform='clipped')


class NormalizedObservationWrapper(Wrapper):
    """Periodic sin / cos encoding of hour, day_type, month and min-max scaling of everything (`citylearn/wrappers.py:39-167`)."""

    def __init__(self, env):
        super().__init__(env)
        self.unwrapped.configure_transforms(observation_transform='normalized')

    @property
    def shared_observations(self) -> List[str]:
        from .schema import PERIODIC_OBSERVATIONS
        out = []
        for o in self.unwrapped.shared_observations:
            out += [f'{o}_cos', f'{o}_sin'] if o in PERIODIC_OBSERVATIONS else [o]
        return out

    @property
    def observation_names(self) -> List[List[str]]:
        u = self.unwrapped
        if u.central_agent:
            return [[n for _, n in u._out_entries]]
        out = [[] for _ in u.spec.buildings]
        for bi, n in u._out_entries:
            out[bi].append(n)
        return out

    @property
    def observation_space(self) -> List[Box]:
        # `estimate_observation_space(normalize=True)`: the unit box (citylearn/building.py:1856-1859)
        return [Box(low=np.zeros(len(n), dtype='float32'), high=np.ones(len(n), dtype='float32'), dtype=np.float32) for n in self.observation_names]


class NormalizedActionWrapper(Wrapper):
    """`step` takes actions in [0, 1]; the kernel maps them to `a * (high - low) + low` (`citylearn/wrappers.py:169-222`)."""

    def __init__(self, env):
        super().__init__(env)
        self.unwrapped.configure_transforms(normalized_actions=True)

    @property
    def action_space(self) -> List[Box]:
        return [Box(low=np.zeros(s.low.size, dtype='float32'), high=np.ones(s.high.size, dtype='float32'), dtype=np.float32)
                for s in self.unwrapped.action_space]


class NormalizedSpaceWrapper(Wrapper):
    """`NormalizedObservationWrapper` + `NormalizedActionWrapper` (`citylearn/wrappers.py:224-238`)."""

    def __init__(self, env):
        super().__init__(NormalizedActionWrapper(NormalizedObservationWrapper(env)))


class StableBaselines3ObservationWrapper(Wrapper):
    """Central-agent observations as ONE flat array per env (`citylearn/wrappers.py:516-545`): with tensors that is the `[E, L]`
    observation tensor itself; reference-shaped lists (`num_envs == 1`) become a float32 vector."""

    def __init__(self, env):
        assert env.unwrapped.central_agent, 'StableBaselines3 wrappers are compatible only when env.central_agent = True.'
        super().__init__(env)

    @property
    def observation_space(self) -> Box:
        return self.env.observation_space[0]

    def observation(self, observations):
        return np.array(observations[0], dtype='float32') if isinstance(observations, list) else observations

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, actions):
        obs, rew, term, trunc, info = self.env.step(actions)
        return self.observation(obs), rew, term, trunc, info


class StableBaselines3ActionWrapper(Wrapper):
    """One flat action vector per env instead of a list of per-agent lists (`citylearn/wrappers.py:547-576`)."""

    def __init__(self, env):
        assert env.unwrapped.central_agent, 'StableBaselines3 wrappers are compatible only when env.central_agent = True.'
        super().__init__(env)

    @property
    def action_space(self) -> Box:
        return self.env.action_space[0]

    def action(self, actions):
        if isinstance(actions, np.ndarray) and actions.ndim == 1 and self.unwrapped.num_envs == 1:
            return [actions.tolist()]
        return actions

    def step(self, actions):
        return self.env.step(self.action(actions))


class StableBaselines3RewardWrapper(Wrapper):
    """Scalar reward for the single agent (`citylearn/wrappers.py:578-601`); tensors keep their `[E, 1]` shape squeezed to `[E]`."""

    def __init__(self, env):
        assert env.unwrapped.central_agent, 'StableBaselines3 wrappers are compatible only when env.central_agent = True.'
        super().__init__(env)

    def reward(self, reward):
        return reward[0] if isinstance(reward, list) else reward.reshape(-1)

    def step(self, actions):
        obs, rew, term, trunc, info = self.env.step(actions)
        return obs, self.reward(rew), term, trunc, info


class StableBaselines3Wrapper(Wrapper):
    """All three Stable-Baselines3 adapters (`citylearn/wrappers.py:603-621`)."""

    def __init__(self, env):
        super().__init__(StableBaselines3ObservationWrapper(StableBaselines3RewardWrapper(StableBaselines3ActionWrapper(env))))
