"""Test-only stand-in for the `gymnasium` package (absent from this image, no network).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference under /root/reference be imported
so that `oracle/make_golden.py` can generate golden traces.  Never imported by the product.
"""
from . import spaces  # noqa: F401


class Env:
    metadata = {}
    render_mode = None

    @property
    def unwrapped(self):
        return self

    def close(self):
        pass


class Wrapper(Env):
    def __init__(self, env):
        self.env = env

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)


class ObservationWrapper(Wrapper):
    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        return self.observation(obs), r, term, trunc, info


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))


class RewardWrapper(Wrapper):
    def step(self, action):
        obs, r, term, trunc, info = self.env.step(action)
        return obs, self.reward(r), term, trunc, info
