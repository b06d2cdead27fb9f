"""citylearn_b200 - B200-native (sm_100a) implementation of CityLearn's per-timestep simulation + reward path.

`CityLearnEnv` keeps the reference's Gymnasium surface; thousands of parallel environments are advanced by one fused
CUDA kernel per step behind a C ABI (`include/citylearn_b200.h`).  Importing this package does not need a GPU; creating
an environment does (there is no CPU fallback).
"""
from .schema import EpisodeTracker, UnknownSchemaError, UnsupportedSchemaError  # noqa: F401
from .data import DataSet  # noqa: F401
from . import reward_function  # noqa: F401

__version__ = '0.1.0'


def __getattr__(name):   # lazy: env imports torch
    if name == 'CityLearnEnv':
        from .env import CityLearnEnv
        return CityLearnEnv
    raise AttributeError(name)
