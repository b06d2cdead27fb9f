"""Reward plug-in surface (mirrors `citylearn/reward_function.py`).

The class names, constructor signatures, `env_metadata` / `central_agent` properties and the
`calculate(observations) -> List[...]` contract are the reference's (`citylearn/reward_function.py:8-386`).
Two things differ, both forced by batching:

* the built-in classes below are *recognised by identity* by `CityLearnEnv` and evaluated inside the fused CUDA step
  kernel (`cl_reward_id` in `include/citylearn_b200.h`); their Python `calculate` is only used when a user
  subclasses them;
* a custom subclass receives observation dicts whose values are `torch.Tensor[E]` (one entry per parallel env) instead
  of Python numbers.  Arithmetic written with operators works unchanged; Python builtins `max/min/abs` on tensors
  must be replaced by `torch.clamp/abs` (or use the helpers in this module).
"""
from __future__ import annotations

from typing import Any, List, Mapping, Tuple

import torch

from .data import ZERO_DIVISION_PLACEHOLDER

__all__ = ['RewardFunction', 'MultiBuildingRewardFunction', 'MARL', 'IndependentSACReward', 'SolarPenaltyReward',
           'ComfortReward', 'SolarPenaltyAndComfortReward', 'Electric_Vehicles_Reward_Function', 'BUILTIN_REWARD_IDS']


def _t(x) -> torch.Tensor:
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x, dtype=torch.float32)


def _total(values: List[Any]):
    out = values[0]
    for v in values[1:]:
        out = out + v
    return out


class RewardFunction:
    r"""Base and default reward: :math:`-\max(e, 0)^{exponent}` per building (`citylearn/reward_function.py:65-88`)."""

    def __init__(self, env_metadata: Mapping[str, Any], exponent: float = None, **kwargs):
        kwargs.pop('charging_constraint_penalty_coefficient', None)
        self.env_metadata = env_metadata
        self.exponent = 1.0 if exponent is None else exponent

    @property
    def env_metadata(self) -> Mapping[str, Any]:
        return self._env_metadata

    @env_metadata.setter
    def env_metadata(self, env_metadata: Mapping[str, Any]):
        self._env_metadata = env_metadata

    @property
    def central_agent(self) -> bool:
        return self.env_metadata['central_agent']

    def reset(self):
        pass

    def _finish(self, reward_list):
        return [_total(reward_list)] if self.central_agent else reward_list

    def calculate(self, observations: List[Mapping[str, Any]]) -> List[Any]:
        reward_list = [-torch.clamp(_t(o['net_electricity_consumption']), min=0.0) ** self.exponent for o in observations]
        return self._finish(reward_list)


class MultiBuildingRewardFunction(RewardFunction):
    """One reward function per building, paired by position (`citylearn/reward_function.py:90-117`)."""

    def __init__(self, env, reward_functions: Mapping[str, RewardFunction]):
        self.reward_functions = reward_functions
        super().__init__(env)

    def calculate(self, observations):
        rewards = []
        for obs, (name, rf) in zip(observations, self.reward_functions.items()):
            if rf is None:
                raise ValueError(f"No reward function for building '{name}'")
            rewards.append(rf.calculate([obs]))   # list of 1-element lists, like the reference
        return rewards

    def reset(self):
        for rf in self.reward_functions.values():
            rf.reset()

    @property
    def env_metadata(self):
        return self._env_metadata

    @env_metadata.setter
    def env_metadata(self, env_metadata):
        self._env_metadata = env_metadata
        for rf in getattr(self, 'reward_functions', {}).values():
            rf.env_metadata = env_metadata


class MARL(RewardFunction):
    """`sign(-e) * 0.01 * e^2 * max(0, district e)` (`citylearn/reward_function.py:132-143`)."""

    def __init__(self, env_metadata: Mapping[str, Any]):
        super().__init__(env_metadata)

    def calculate(self, observations):
        e = [_t(o['net_electricity_consumption']).double() for o in observations]
        district = torch.clamp(_total(e), min=0.0)
        reward_list = [torch.sign(-v) * 0.01 * v ** 2 * district for v in e]
        return self._finish(reward_list)


class IndependentSACReward(RewardFunction):
    """`min(-e, 0)` - the reference's `v * -1 ** 3` parses as `-v` (`citylearn/reward_function.py:159-168`)."""

    def __init__(self, env_metadata: Mapping[str, Any]):
        super().__init__(env_metadata)

    def calculate(self, observations):
        return self._finish([torch.clamp(-_t(o['net_electricity_consumption']), max=0.0) for o in observations])


class SolarPenaltyReward(RewardFunction):
    """Penalise consumption when storage is empty and export when it is full (`citylearn/reward_function.py:189-214`)."""

    def __init__(self, env_metadata: Mapping[str, Any]):
        super().__init__(env_metadata)

    def calculate(self, observations):
        reward_list = []
        for o, m in zip(observations, self.env_metadata['buildings']):
            e = _t(o['net_electricity_consumption'])
            reward = torch.zeros_like(e)
            for dev in ('cooling_storage', 'heating_storage', 'dhw_storage', 'electrical_storage'):
                if m[dev]['capacity'] > ZERO_DIVISION_PLACEHOLDER:
                    soc = _t(o.get(f'{dev}_soc', 0.0))
                    reward = reward - (1.0 + torch.sign(e) * soc) * torch.abs(e)
            reward_list.append(reward)
        return self._finish(reward_list)


class ComfortReward(RewardFunction):
    """Indoor-temperature comfort band reward (`citylearn/reward_function.py:269-334`)."""

    def __init__(self, env_metadata: Mapping[str, Any], band: float = None, lower_exponent: float = None, higher_exponent: float = None):
        super().__init__(env_metadata)
        self.band = band
        self.lower_exponent = 2.0 if lower_exponent is None else lower_exponent
        self.higher_exponent = 2.0 if higher_exponent is None else higher_exponent

    def _one(self, o):
        lo, hi = self.lower_exponent, self.higher_exponent
        heating = _t(o.get('heating_demand', 0.0)) > _t(o.get('cooling_demand', 0.0))
        mode = _t(o['hvac_mode'])
        T = _t(o['indoor_dry_bulb_temperature'])
        csp = _t(o['indoor_dry_bulb_temperature_cooling_set_point'])
        hsp = _t(o['indoor_dry_bulb_temperature_heating_set_point'])
        band = _t(o['comfort_band']) if self.band is None else _t(self.band)
        w = torch.where
        # modes 1 (cooling) and 2 (heating): one set point
        sp = w(mode == 1, csp, hsp)
        delta = torch.abs(T - sp)
        r12 = w(T < sp - band, -delta ** w(mode == 2, _t(lo), _t(hi)),
                w(T < sp, w(heating, torch.zeros_like(delta), -delta),
                  w(T <= sp + band, w(heating, -delta, torch.zeros_like(delta)), -delta ** w(heating, _t(hi), _t(lo)))))
        # other modes: dead band between the two set points
        cd, hd = T - csp, T - hsp
        r03 = w(T < hsp - band, -torch.abs(hd) ** w(~heating, _t(hi), _t(lo)),
                w(T < hsp, -torch.abs(hd),
                  w(T <= csp, torch.zeros_like(T),
                    w(T < csp + band, -torch.abs(cd), -torch.abs(cd) ** w(heating, _t(hi), _t(lo))))))
        out = w((mode == 1) | (mode == 2), r12, r03)
        return w(torch.isnan(T), torch.full_like(out, float('nan')), out)

    def calculate(self, observations):
        return self._finish([self._one(o) for o in observations])


class SolarPenaltyAndComfortReward(RewardFunction):
    """Weighted sum of the two rewards above (`citylearn/reward_function.py:336-386`)."""

    def __init__(self, env_metadata: Mapping[str, Any], band: float = None, lower_exponent: float = None,
                 higher_exponent: float = None, coefficients: Tuple = None):
        self._functions: List[RewardFunction] = [
            SolarPenaltyReward(env_metadata),
            ComfortReward(env_metadata, band=band, lower_exponent=lower_exponent, higher_exponent=higher_exponent)]
        super().__init__(env_metadata)
        coefficients = [1.0] * len(self._functions) if coefficients is None else coefficients
        assert len(coefficients) == len(self._functions), f'{type(self).__name__} needs {len(self._functions)} coefficients.'
        self.coefficients = coefficients

    @property
    def env_metadata(self):
        return self._env_metadata

    @env_metadata.setter
    def env_metadata(self, env_metadata):
        self._env_metadata = env_metadata
        for f in getattr(self, '_functions', []):
            f.env_metadata = env_metadata

    def calculate(self, observations):
        parts = [f.calculate(observations) for f in self._functions]
        return [_total([p[i].double() * c for p, c in zip(parts, self.coefficients)]) for i in range(len(parts[0]))]


class Electric_Vehicles_Reward_Function(MARL):
    """`citylearn.reward_function.Electric_Vehicles_Reward_Function` (reward_function.py:389-523): the MARL reward only scales the
    penalty / bonus terms of a building's chargers (battery limits, reachability and closeness of the required departure SOC,
    self-consumption / self-production); a building without chargers is rewarded 0.  Evaluated inside the step kernel (default weights;
    custom weights would need the Python path, which this class does not provide: the per-charger dictionaries live on the device)."""

    def __init__(self, env_metadata: Mapping[str, Any] = None, weights: Mapping[str, float] = None,
                 charging_constraint_penalty_coefficient: float = None, **kwargs):
        super().__init__(env_metadata, **kwargs)
        # reward_function.py:22-25, 50-58: multiplies a building's `charging_constraint_violation_kwh` (default 1.0)
        self.charging_constraint_penalty_coefficient = 1.0 if charging_constraint_penalty_coefficient is None else float(charging_constraint_penalty_coefficient)
        if weights:
            raise NotImplementedError('Electric_Vehicles_Reward_Function: custom weights are not supported (the fused kernel uses the defaults)')
        self.weights = {"no_car_charging": -5.0, "battery_limits": -2.0, "soc_impossible": -10.0, "soc_under": -5.0, "close_soc": 10.0,
                        "self_ev_consumption": 5.0, "extra_self_production": 5.0}

    def calculate(self, observations):
        raise NotImplementedError('Electric_Vehicles_Reward_Function is evaluated by the step kernel (cl_reward_id 6)')


# class -> cl_reward_id (include/citylearn_b200.h); identity match only, subclasses go through the Python path
BUILTIN_REWARD_IDS = {
    RewardFunction: 0, MARL: 1, IndependentSACReward: 2, SolarPenaltyReward: 3, ComfortReward: 4, SolarPenaltyAndComfortReward: 5,
    Electric_Vehicles_Reward_Function: 6,
}
