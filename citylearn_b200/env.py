"""`CityLearnEnv`: the reference's Gymnasium surface over the fused CUDA step path.

Drop-in contract (reference `citylearn/citylearn.py`):
  * `CityLearnEnv(schema, **overrides)` with the same override keywords (`:133-142`, precedence `:2006-2051`);
  * `reset(seed=None, options=None) -> (observations, info)` (`:1829-1886`);
  * `step(actions) -> (observations, reward, terminated, truncated, info)` (`:978-1056`);
  * `observation_space`, `action_space`, `observation_names`, `action_names`, `buildings`, `time_step`, `time_steps`,
    `terminated`, `truncated`, `episode_rewards`, `episode_tracker`, `get_metadata()`, `unwrapped` (`:384-538, :946-953`).

With `num_envs=1` and list actions the return values have the reference's shapes (lists of lists).  New keywords:
`num_envs` (parallel environments held as CUDA tensors), `device`, `precision` ('fp64': the reference's own
float64-intermediate / float32-storage arithmetic, bit-exact physics; 'fp32': float arithmetic), `stale_observations`
(True = reference parity: action-dependent entries of the observation returned by `step` are the zero-initialised
values at t+1; False = they carry the post-action values of step t).

Batched mode (`num_envs > 1`, or tensor / ndarray actions): `actions` is `[E, sum(A_b)]` (district vector: buildings in
order, active actions in schema order) or a list of per-building `[E, A_b]`; `step` returns `obs [E, L]`, `reward [E, B]`
(or `[E, 1]` with a central agent) as CUDA tensors that are overwritten by the next step, and Python bools for
terminated / truncated (all envs advance in lock-step).
"""
from __future__ import annotations

import importlib
import os
from typing import Any, Dict, List, Mapping, Optional, Tuple, Union

import numpy as np
import torch

from . import _native
from . import reward_function as rf_mod
from . import schema as S
from .spaces import Box

_REFERENCE_REWARD_MODULES = ('citylearn.reward_function', 'citylearn_b200.reward_function')


def _resolve_reward_class(spec_type):
    if isinstance(spec_type, type):
        # a reference class object (citylearn.reward_function.X) maps to the class of the same name here
        if spec_type.__module__ in _REFERENCE_REWARD_MODULES and hasattr(rf_mod, spec_type.__name__):
            return getattr(rf_mod, spec_type.__name__)
        return spec_type
    module_name, _, class_name = str(spec_type).rpartition('.')
    if module_name in _REFERENCE_REWARD_MODULES and hasattr(rf_mod, class_name):
        return getattr(rf_mod, class_name)
    return getattr(importlib.import_module(module_name), class_name)


class BuildingProxy:
    """Light stand-in for `citylearn.building.Building`: the attributes wrappers / agents read."""

    def __init__(self, env: 'CityLearnEnv', spec: S.BuildingSpec):
        self._env = env
        self._spec = spec
        self.name = spec.name
        self.observation_metadata = dict(spec.observation_metadata)
        self.action_metadata = dict(spec.action_metadata)
        self.observation_space = Box(low=spec.observation_low, high=spec.observation_high, dtype=np.float32)
        self.action_space = Box(low=spec.action_low, high=spec.action_high, dtype=np.float32)
        self.time_step_ratio = spec.time_step_ratio
        self.seconds_per_time_step = spec.seconds_per_time_step
        self._norm_limits = None

    @property
    def active_observations(self) -> List[str]:
        return self._spec.active_observations

    @property
    def active_actions(self) -> List[str]:
        return self._spec.active_actions

    # ---- what the reference's own wrappers / agents call on `env.unwrapped.buildings[i]` (citylearn/wrappers.py:90-165, 193-204) ----
    @property
    def index(self) -> int:
        return self._env.spec.buildings.index(self._spec)

    def _current_values(self) -> Dict[str, float]:
        """name -> value of this building's ACTIVE observations at the env's current time step (what `Building.observations()` walks,
        citylearn/building.py:1115-1160): series values from the table, action-dependent values from the env's observation row."""
        env = self._env
        if env.num_envs != 1:
            raise RuntimeError('BuildingProxy.observations() mirrors the single-env reference API; batched envs read `env.observations`')
        if env._observation_transform is not None:
            raise RuntimeError('BuildingProxy.observations() needs the raw observation row (the env has a fused observation transform)')
        bi = self.index
        if env._proxy_layout is None:
            entries, desc = S.observation_layout(env.spec, False, env.stale_observations)
            env._proxy_layout = (entries, desc, {key: j for j, key in enumerate(env._entries)})
        entries, desc, pos = env._proxy_layout
        row = env.observations
        flat = row[0] if env.central_agent else [v for r in row for v in r]
        t = env.time_step
        trow = env.spec.table[int(env.episode_tracker.episode_start_time_step) + t]
        out: Dict[str, float] = {}
        for (b2, name), d in zip(entries, desc):
            if b2 != bi:
                continue
            if d[0] == S.OBS_TS:
                out[name] = float(trow[d[1]])
            elif d[0] == S.OBS_OUTAGE:
                out[name] = float(env._outage[bi, t])
            else:
                out[name] = float(flat[pos[(bi, name)]])
        return out

    def estimate_observation_space_limits(self, include_all: bool = None, periodic_normalization: bool = None):
        return S.estimate_observation_space_limits(self._spec, self._env.spec, include_all=bool(include_all), periodic_normalization=bool(periodic_normalization))

    def estimate_observation_space(self, include_all: bool = None, normalize: bool = None) -> Box:
        """citylearn/building.py:1836-1865."""
        if normalize:
            lo, _ = self.estimate_observation_space_limits(include_all, True)
            return Box(low=np.zeros(len(lo), dtype='float32'), high=np.ones(len(lo), dtype='float32'), dtype=np.float32)
        lo, hi = self.estimate_observation_space_limits(include_all, False)
        return Box(low=np.array(list(lo.values()), dtype='float32'), high=np.array(list(hi.values()), dtype='float32'), dtype=np.float32)

    def observations(self, include_all: bool = None, normalize: bool = None, periodic_normalization: bool = None,
                     check_limits: bool = None) -> Mapping[str, float]:
        """Observations at the current time step as the reference's `Building.observations` returns them (citylearn/building.py:1115-1219):
        active observations in schema order, optionally `<name>_cos`, `<name>_sin` for the periodic ones (in that order) and min-max
        normalisation with the `include_all=True` limits (`x_min == x_max` gives 0, citylearn/preprocessing.py:139-152)."""
        if include_all:
            raise NotImplementedError('include_all=True (the 67-key reward dictionary) is served to reward functions directly; wrappers use active observations')
        obs = self._current_values()
        if periodic_normalization:
            out: Dict[str, float] = {}
            for k, v in obs.items():
                if k in S.PERIODIC_OBSERVATIONS:
                    x = 2 * np.pi * v / S.PERIODIC_OBSERVATIONS[k]
                    out[f'{k}_cos'] = float(np.cos(x))
                    out[f'{k}_sin'] = float(np.sin(x))
                else:
                    out[k] = v
            obs = out
        if normalize:
            if self._norm_limits is None:
                self._norm_limits = S.estimate_observation_space_limits(self._spec, self._env.spec, include_all=True, periodic_normalization=True)
            lo, hi = self._norm_limits
            for k, v in obs.items():
                x_min, x_max = lo[k], hi[k]
                obs[k] = 0 if x_min == x_max else (v - x_min) / (x_max - x_min)
        return obs

    def get_metadata(self) -> Mapping[str, Any]:
        dv = self._spec.devices
        s = self._spec.series
        n_years = max(1, self._env.time_steps * self.seconds_per_time_step / (8760 * 3600))

        def device(d):
            out = {k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in d.items() if k not in ('type', 'class', 'absent')}
            if 'efficiency' in d and d['type'] in ('StorageTank', 'Battery'):
                out['round_trip_efficiency'] = d['efficiency'] ** 0.5
            return out

        return {
            'name': self.name, 'observation_metadata': self.observation_metadata, 'action_metadata': self.action_metadata,
            **{k: device(dv[k]) for k in ('cooling_device', 'heating_device', 'dhw_device', 'cooling_storage', 'heating_storage',
                                          'dhw_storage', 'electrical_storage', 'pv')},
            'annual_cooling_demand_estimate': float(s['cooling_demand'].sum() / n_years),
            'annual_heating_demand_estimate': float(s['heating_demand'].sum() / n_years),
            'annual_dhw_demand_estimate': float(s['dhw_demand'].sum() / n_years),
            'annual_non_shiftable_load_estimate': float(s['non_shiftable_load'].sum() / n_years),
            'annual_solar_generation_estimate': float(S.pv_generation(self._spec, s['solar_generation']).sum() / n_years),
        }


class CityLearnEnv:
    metadata: Dict[str, Any] = {}
    render_mode = None

    def __new__(cls, *args, devices=None, **kwargs):
        # `devices=[...]` with more than one entry: the envs are sharded over the devices of THIS process; the object
        # returned is a `distributed.DeviceShardedEnv` holding one CityLearnEnv per device
        if cls is CityLearnEnv and devices is not None and len(list(devices)) > 1:
            from .distributed import DeviceShardedEnv
            return DeviceShardedEnv(*args, devices=list(devices), **kwargs)
        return super().__new__(cls)

    def __init__(self, schema, num_envs: int = 1, device: Union[str, torch.device, None] = None, precision: str = 'fp64',
                 stale_observations: bool = True, track_episode_rewards: Optional[bool] = None, debug_trace: bool = False,
                 record_history: Optional[bool] = None, history_env: int = 0, observation_transform: Optional[str] = None,
                 normalized_actions: bool = False, track_kpis: bool = False, auto_reset: bool = False, devices=None, **kwargs):
        if devices is not None and device is None and len(list(devices)) == 1:
            device = list(devices)[0]
        self.spec = S.load(schema, **kwargs)
        self.schema = self.spec.schema
        self.num_envs = int(num_envs)
        if self.num_envs < 1:
            raise ValueError('num_envs must be >= 1')
        if not torch.cuda.is_available():
            raise RuntimeError('citylearn_b200 needs a CUDA device (B200 / sm_100a); there is no CPU fallback.')
        self.device = torch.device('cuda' if device is None else device)
        if self.device.type != 'cuda':
            raise RuntimeError("citylearn_b200 runs on CUDA devices only (device='cuda[:i]'); there is no CPU fallback.")
        if self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        self.precision = precision
        self.stale_observations = bool(stale_observations)
        # batched episode end: opt-in Gymnasium-VectorEnv "same-step" auto-reset - the step that ends the episode resets
        # every env (they advance in lock-step) and returns the first observation of the next episode, the last one of the finished
        # episode in info['final_observation'].  Never at num_envs == 1: the reference leaves reset() to the caller (citylearn.py:995-997)
        self.auto_reset = bool(auto_reset)
        if self.auto_reset and self.num_envs == 1:
            raise ValueError('auto_reset is a batched-env option; with num_envs == 1 the caller resets, like the reference')
        spec = self.spec
        self.central_agent = spec.central_agent
        self.shared_observations = spec.shared_observations
        self.random_seed = spec.random_seed
        self.seconds_per_time_step = spec.seconds_per_time_step
        self.time_step_ratio = spec.time_step_ratio
        self.episode_time_steps = spec.episode_time_steps
        self.rolling_episode_split = spec.rolling_episode_split
        self.random_episode_split = spec.random_episode_split
        self.episode_tracker = S.EpisodeTracker(spec.simulation_start_time_step, spec.simulation_end_time_step)
        self.buildings = [BuildingProxy(self, b) for b in spec.buildings]
        # `_entries` is the raw observation row (what `observation_names` / `observation_space` describe); `_out_entries` is the
        # row the kernels write, which differs under a fused observation wrapper (citylearn_b200/wrappers.py)
        self._entries, self._raw_desc = S.observation_layout(spec, self.central_agent, self.stale_observations)
        # names / space order (differs from the value order only for buildings with charging constraints, like the reference's)
        self._name_entries = S.observation_layout(spec, self.central_agent, self.stale_observations, names_order=True)[0]
        self._observation_transform = observation_transform
        self._normalized_actions = bool(normalized_actions)
        self._sizes_act = [len(b.active_actions) for b in spec.buildings]
        self._track = (self.num_envs == 1) if track_episode_rewards is None else bool(track_episode_rewards)
        # per-step history of ONE env for evaluate() (the reference keeps full series for its single env)
        # districts with electric vehicles / chargers / washing machines: the fused step only - reference-parity
        # observations, built-in rewards, no KPI history (the reference's evaluate() adds charger cost functions this build does not have)
        evd = spec.ev or {}
        self._has_ev = bool(len(evd.get('chargers', ())) or len(evd.get('wms', ())))
        if self._has_ev:
            if not self.stale_observations:
                raise NotImplementedError('districts with electric vehicles / washing machines need stale_observations=True')
            if record_history or track_kpis or debug_trace:
                raise NotImplementedError('record_history / track_kpis / debug_trace are not available for districts with electric vehicles / washing machines')
            if self.central_agent:
                # (loader and oracle are pinned against the reference for a central agent - tests/golden/ev_cpu - but the kernel's summed
                #  reward of this instantiation has not run on hardware yet: refuse rather than return unverified numbers)
                raise NotImplementedError('central_agent=True is not available for districts with electric vehicles / washing machines yet')
            record_history = False
        self._record = (self.num_envs == 1) if record_history is None else bool(record_history)
        self._history_env = int(history_env)
        if not 0 <= self._history_env < self.num_envs:
            raise ValueError('history_env out of range')
        # online KPI accumulators for every env: fed from the per-step trace by a small second kernel
        self._track_kpis = bool(track_kpis)
        debug_trace = debug_trace or self._record        # (track_kpis adds the trace only where the accumulators are not fused, _build_native)
        # reward function (citylearn/citylearn.py:2100-2163)
        self.reward_function = self._make_reward_function()
        rid, rparams = self._fused_reward()
        self._reward_id = rid
        if any(b.observation_value_order is not None for b in spec.buildings) and observation_transform is not None:
            raise NotImplementedError('observation wrappers on districts with charging constraints: the reference pairs the space limits with '
                                      'the values by position although their orders differ (building.py:1146-1154); not reproduced')
        if self._has_ev and rid < 0:
            raise NotImplementedError('districts with electric vehicles / washing machines need a built-in reward function (evaluated in the step kernel)')
        self._reward_dim = 1 if self.central_agent else spec.n_buildings
        self._reward_params = rparams
        self._debug_trace = debug_trace
        self._h = None
        self._proxy_layout = None
        self._build_native()
        self._table_dev = None
        self.time_step = 0
        self._episode_rewards: List[Mapping[str, Any]] = []
        self._rsum = self._rmin = self._rmax = None
        self.reset()
        self.episode_tracker.reset_episode_index()       # citylearn/citylearn.py:237-240
        self.reward_function.env_metadata = self.get_metadata()
        self._episode_rewards = []

    def _build_native(self):
        """(Re)create the device-side district and the output buffers for the current observation / action transforms."""
        spec, rid, precision = self.spec, self._reward_id, self.precision
        debug_trace = self._debug_trace
        if self._observation_transform is None:
            self._out_entries, self._desc, transforms = self._entries, self._raw_desc, None
        else:
            self._out_entries, self._desc, transforms = S.transformed_observation_layout(spec, self._entries, self._raw_desc, self._observation_transform)
        self._obs_dim = len(self._out_entries)
        counts = np.bincount(np.fromiter((bi for bi, _ in self._out_entries), dtype=np.int64, count=self._obs_dim), minlength=spec.n_buildings)
        self._sizes_obs = [int(c) for c in counts]
        if self._h is not None:
            self._h.close()
        with torch.cuda.device(self.device):
            self._h = _native.Handle(spec, self.num_envs, self._desc, self.central_agent, rid, self._reward_params, precision, self.stale_observations)
            if transforms is not None or self._normalized_actions:
                lo = np.array([v for b in spec.buildings for v in b.action_low], dtype='float32')
                hi = np.array([v for b in spec.buildings for v in b.action_high], dtype='float32')
                self._h.set_transforms(transforms, (hi - lo) if self._normalized_actions else None, lo if self._normalized_actions else None)
            self._kpi_fused = False
            if self._track_kpis:
                self._h.kpi_enable(True)
                self._kpi_fused = self._h.kpi_fused()        # accumulated inside the step kernel: no trace, survives rollout()
                debug_trace = debug_trace or not self._kpi_fused
            self._kpi_valid = True
            E = self.num_envs
            # observations, rewards and the shared observation row live in one allocation [obs E*L | reward E*R | row L] so that the
            # host path needs ONE device->host copy per step: head (obs + reward) or tail (reward + the row every env shares)
            nL, nR = E * self._obs_dim, E * self._reward_dim
            self._out = torch.zeros(nL + nR + self._obs_dim, dtype=torch.float32, device=self.device)
            self._obs = self._out[:nL].view(E, self._obs_dim)
            self._reward = self._out[nL:nL + nR].view(E, self._reward_dim)
            self._row = self._out[nL + nR:]
            self._obs_current = True            # False after a shared-row host step: self._obs was not written
            self._district = torch.zeros((E, 3), dtype=torch.float32, device=self.device)
            self._trace = (torch.zeros((E, spec.n_buildings, S.NDYN), dtype=torch.float32, device=self.device)
                           if (rid < 0 or debug_trace) else None)
            self._act = torch.zeros((E, max(spec.action_dim, 1)), dtype=torch.float32, device=self.device)
            # host -> device action staging: a small ring of pinned buffers, each guarded by a CUDA event, because the H2D copy is
            # asynchronous - the host must not refill a buffer before the copy that reads it has executed
            self._stage_n = 4
            self._stage_pinned = [torch.zeros((E, max(spec.action_dim, 1)), dtype=torch.float32).pin_memory() for _ in range(self._stage_n)]
            self._stage_host = [t.numpy() for t in self._stage_pinned]      # same memory; filled with np.copyto (single-threaded memcpy)
            self._stage_event = [torch.cuda.Event() for _ in range(self._stage_n)]
            self._stage_used = [False] * self._stage_n
            self._stage_i = 0
            self._out_pinned = torch.zeros(nL + nR + self._obs_dim, dtype=torch.float32).pin_memory()
            self._obs_pinned = self._out_pinned[:nL].view(E, self._obs_dim)
            self._reward_pinned = self._out_pinned[nL:nL + nR].view(E, self._reward_dim)
            self._obs_host, self._reward_host = self._obs_pinned.numpy(), self._reward_pinned.numpy()
            self._row_host = self._out_pinned[nL + nR:].numpy()
            self._row_view = np.broadcast_to(self._row_host, (E, self._obs_dim))
            # single-call host step (cl_step_host): possible when nothing but the fused kernel has to run per step
            self._host_fast = (rid >= 0 and self._trace is None and not self._track and not self._record and not self.auto_reset
                               and (not self._track_kpis or self._kpi_fused))
            self._host_ptrs = {'act': self._act.data_ptr(), 'obs': self._obs.data_ptr(), 'reward': self._reward.data_ptr(), 'row': self._row.data_ptr(),
                               'district': self._district.data_ptr(), 'obs_host': self._obs_pinned.data_ptr(), 'reward_host': self._reward_pinned.data_ptr()}
            self._pinned_action_buffers: List[torch.Tensor] = []
            # in-place host step: actions read from / rewards + row written to page-locked host memory by the kernels themselves
            # cl_step_host flags (include/citylearn_b200.h): 1 read actions in place, 2 write results in place, 4 polled completion flag
            self._host_in_place = int(os.environ.get('CL_B200_HOST_MODE', '3'))
            self._reward_current = True
            self._roll = None                   # buffers of rollout_host, keyed by K

    def configure_transforms(self, observation_transform: Optional[str] = 'unchanged', normalized_actions: Optional[bool] = None):
        """Fuse wrapper semantics into the kernels (used by `citylearn_b200.wrappers`): `observation_transform` in
        {None, 'normalized', 'clipped'}; `normalized_actions`: step() takes fractions of the action range.  Rebuilds the
        device-side district, i.e. starts a fresh episode like a reset."""
        if observation_transform not in ('unchanged', None) and any(b.observation_value_order is not None for b in self.spec.buildings):
            raise NotImplementedError('observation wrappers are not available for districts with charging constraints')
        if observation_transform != 'unchanged':
            self._observation_transform = observation_transform
        if normalized_actions is not None:
            self._normalized_actions = bool(normalized_actions)
        self._build_native()
        self.episode_tracker.reset_episode_index()
        self.reset()
        self.episode_tracker.reset_episode_index()
        self._episode_rewards = []

    # ---------------------------------------------------------------------------------------------
    # reward plumbing
    # ---------------------------------------------------------------------------------------------
    def _make_reward_function(self):
        spec = self.spec
        rt, attrs = spec.reward_type, spec.reward_attributes
        if isinstance(rt, dict):   # per-building reward functions (citylearn/citylearn.py:2106-2141)
            default_type = rt.get('default') or (next(iter(rt.values())) if rt else None)
            default_attrs = (attrs or {}).get('default')
            if default_attrs is None and attrs:
                default_attrs = next(iter(attrs.values()))
            fns = {}
            for b in spec.buildings:
                r_type = rt.get(b.name, default_type)
                if r_type is None:
                    raise ValueError(f"No reward function defined for building '{b.name}' and no default provided")
                fns[b.name] = _resolve_reward_class(r_type)(None, **((attrs or {}).get(b.name, default_attrs) or {}))
            return rf_mod.MultiBuildingRewardFunction(None, fns)
        if isinstance(rt, rf_mod.RewardFunction):
            return rt
        return _resolve_reward_class(rt)(None, **(attrs or {}))

    def _fused_reward(self) -> Tuple[int, List[float]]:
        r = self.reward_function
        rid = rf_mod.BUILTIN_REWARD_IDS.get(type(r), -1)
        if rid == 0:
            return rid, [float(r.exponent)]
        if rid == 6:
            return rid, [float(r.charging_constraint_penalty_coefficient)]
        if rid == 4:
            return rid, [float('nan') if r.band is None else float(r.band), float(r.lower_exponent), float(r.higher_exponent)]
        if rid == 5:
            c = r._functions[1]
            return rid, [float('nan') if c.band is None else float(c.band), float(c.lower_exponent), float(c.higher_exponent),
                         float(r.coefficients[0]), float(r.coefficients[1])]
        return rid, []

    def _reward_observations(self) -> List[Dict[str, torch.Tensor]]:
        """Per-building dicts of `Tensor[E]` at step t for Python reward functions (citylearn/citylearn.py:1022)."""
        spec = self.spec
        if self._table_dev is None:
            self._table_dev = torch.as_tensor(spec.table, device=self.device)
        rows = self._start_dev.long() + (self.time_step)
        out = []
        for bi, b in enumerate(spec.buildings):
            d: Dict[str, torch.Tensor] = {}
            for key, col in spec.columns.items():
                if isinstance(key, tuple) and len(key) == 2 and key[0] == bi and isinstance(key[1], str) and not key[1].startswith('C_'):
                    d[key[1]] = self._table_dev[rows, col]
            for name, slot in S.DYN.items():
                d[name] = self._trace[:, bi, slot]
            d['power_outage'] = torch.full((self.num_envs,), float(self._outage[bi, self.time_step]), device=self.device)
            out.append(d)
        return out

    # ---------------------------------------------------------------------------------------------
    # Gymnasium surface
    # ---------------------------------------------------------------------------------------------
    @property
    def unwrapped(self):
        return self

    @property
    def time_steps(self) -> int:
        return self.episode_tracker.episode_time_steps

    @property
    def terminated(self) -> bool:
        return self.time_step == self.time_steps - 1     # citylearn/citylearn.py:372-376

    @property
    def truncated(self) -> bool:
        return False

    @property
    def episode_rewards(self):
        return self._episode_rewards

    @property
    def observation_space(self) -> List[Box]:
        if self.central_agent:
            lo, hi = [], []
            flat_lo = {(bi, n): v for bi, b in enumerate(self.spec.buildings) for n, v in zip(b.active_observations, b.observation_low)}
            flat_hi = {(bi, n): v for bi, b in enumerate(self.spec.buildings) for n, v in zip(b.active_observations, b.observation_high)}
            for key in self._name_entries:
                lo.append(flat_lo[key])
                hi.append(flat_hi[key])
            return [Box(low=np.array(lo, dtype='float32'), high=np.array(hi, dtype='float32'), dtype=np.float32)]
        return [b.observation_space for b in self.buildings]

    @property
    def action_space(self) -> List[Box]:
        if self.central_agent:
            lo = [v for b in self.spec.buildings for v in b.action_low]
            hi = [v for b in self.spec.buildings for v in b.action_high]
            return [Box(low=np.array(lo, dtype='float32'), high=np.array(hi, dtype='float32'), dtype=np.float32)]
        return [b.action_space for b in self.buildings]

    @property
    def observation_names(self) -> List[List[str]]:
        if self.central_agent:
            return [[n for _, n in self._entries]]
        # the keys of `Building.observations()` (citylearn.py:498-520): the order of the VALUES; the space follows `active_observations`
        return [list(b.observation_value_order or b.active_observations) for b in self.spec.buildings]

    @property
    def action_names(self) -> List[List[str]]:
        if self.central_agent:
            return [[n for b in self.spec.buildings for n in b.active_actions]]
        return [list(b.active_actions) for b in self.spec.buildings]

    @property
    def observations(self):
        """Current observation: reference-shaped lists for num_envs == 1, else the `[E, L]` CUDA tensor."""
        if not self._obs_current and self.shared_observation_row:
            # the last step(s) did not materialise the [E, L] slab (shared-row host path, rollout without `obs`): every env's row is
            # the shared row of the current time step
            with torch.cuda.device(self.device):
                self._h.obs_rows(self.time_step, 1, self._row.data_ptr(), self._stream())
                self._obs.copy_(self._row.expand(self.num_envs, self._obs_dim))
            self._obs_current = True
        return self._shape_obs(self._obs) if self.num_envs == 1 else self._obs

    def get_metadata(self) -> Mapping[str, Any]:
        return {
            'random_seed': self.random_seed, 'simulation_time_steps': self.episode_tracker.simulation_time_steps,
            'seconds_per_time_step': self.seconds_per_time_step, 'time_step_ratio': self.time_step_ratio,
            'buildings': [b.get_metadata() for b in self.buildings], 'central_agent': self.central_agent,
            'shared_observations': self.shared_observations, 'num_envs': self.num_envs,
        }

    def get_info(self) -> Mapping[Any, Any]:
        return {}

    def close(self):
        with torch.cuda.device(self.device):
            self._h.close()

    # ---------------------------------------------------------------------------------------------
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _shape_obs(self, obs: torch.Tensor):
        row = obs[0].tolist()
        if self.central_agent:
            return [row]
        out, o = [], 0
        for n in self._sizes_obs:
            out.append(row[o:o + n])
            o += n
        return out

    def reset(self, seed: int = None, options: Mapping[str, Any] = None, mask=None):
        """Start the next episode (citylearn/citylearn.py:1829-1886). `options={'episode_start': Tensor[E] int32}` gives every
        env its own window start (same length) instead of the tracker's.  `mask` (bool [E]): which envs to reset - the envs of a
        handle share ONE time step (every env ends its episode on the same step, also with per-env windows), so only the full mask
        is meaningful; a partial mask raises."""
        if mask is not None:
            m = torch.as_tensor(mask).reshape(-1).bool()
            if m.numel() != self.num_envs:
                raise ValueError(f'mask must have one entry per env ({self.num_envs})')
            if not bool(m.all()):
                raise NotImplementedError('partial reset: the envs of a CityLearnEnv advance in lock-step and terminate together; reset all of them '
                                          '(mask=None), or hold groups that must restart independently in separate CityLearnEnv objects')
        if seed is not None:
            self.random_seed = seed
        ets = self.episode_time_steps if self.episode_time_steps is not None else self.episode_tracker.simulation_time_steps
        self.episode_tracker.next_episode(ets, self.rolling_episode_split, self.random_episode_split, self.random_seed)
        start = self.episode_tracker.episode_start_time_step
        T = self.episode_tracker.episode_time_steps
        self._outage = S.outage_signals(self.spec, T, start)
        self.time_step = 0
        with torch.cuda.device(self.device):
            stream = self._stream()
            self._h.set_outage(self._outage if any(b.simulate_power_outage for b in self.spec.buildings) else None, stream)
            per_env = None if not options else options.get('episode_start')
            if per_env is not None:
                starts = torch.as_tensor(per_env).to(torch.int64).reshape(-1).cpu()
                if starts.shape != (self.num_envs,):
                    raise ValueError(f"options['episode_start'] must hold one table row per env ({self.num_envs}), got {tuple(starts.shape)}")
                lo, hi = int(starts.min()), int(starts.max())
                first, last = self.episode_tracker.simulation_start_time_step, self.episode_tracker.simulation_end_time_step
                if lo < first or hi + T - 1 > last or hi + T > self.spec.table.shape[0]:
                    raise ValueError(f"options['episode_start']: every window [start, start + {T}) must lie inside the simulation period "
                                     f'[{first}, {last}] (got starts in [{lo}, {hi}])')
                if self._outage.any():
                    raise ValueError('per-env episode starts cannot be combined with power-outage signals (one signal per episode window)')
                self._start_dev = starts.to(device=self.device, dtype=torch.int32).contiguous()
                self._h.reset(self._start_dev.data_ptr(), 0, T, self._obs.data_ptr(), stream)
            else:
                self._start_dev = torch.full((self.num_envs,), start, dtype=torch.int32, device=self.device)
                self._h.reset(None, start, T, self._obs.data_ptr(), stream)
        self.reward_function.reset()
        self._rsum = self._rmin = self._rmax = None
        if self._record:
            self._hist_dyn = torch.zeros((T - 1, self.spec.n_buildings, S.NDYN), dtype=torch.float32, device=self.device)
            self._hist_district = torch.zeros((T - 1, 3), dtype=torch.float32, device=self.device)
            self._hist_valid = True
        self._kpi_valid = True
        self._obs_current = True
        self._uniform_start = per_env is None
        if self.num_envs == 1:
            return self._shape_obs(self._obs), self.get_info()
        return self._obs, self.get_info()

    def _parse_actions(self, actions) -> Tuple[torch.Tensor, bool]:
        """-> (device tensor [E, A], reference_shaped)."""
        E, A = self.num_envs, self.spec.action_dim
        if isinstance(actions, torch.Tensor):
            a = actions.reshape(E, A)
            if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous():
                a = a.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
            return a, False
        if isinstance(actions, np.ndarray):
            # plain memcpy into the pinned staging buffer: a torch CPU copy_ would fan out over the intra-op thread pool, and
            # the spinning pool threads can exhaust a container's CPU quota (observed: 70 ms cgroup throttling stalls)
            self._upload(lambda host: np.copyto(host[:, :A], actions.reshape(E, A), casting='same_kind'))
            return self._act, False
        actions = list(actions)
        if len(actions) and isinstance(actions[0], torch.Tensor):      # per-building [E, A_b]
            a = torch.cat([x.reshape(E, -1).to(self.device, torch.float32) for x in actions], dim=1).contiguous()
            assert a.shape[1] == A
            return a, False
        # reference-style nested lists (citylearn/citylearn.py:1063-1134)
        if E != 1:
            raise ValueError('nested-list actions are only accepted with num_envs == 1; pass an [E, A] tensor or ndarray')
        if self.central_agent:
            flat = [float(v) for v in actions[0]]
            assert len(flat) == A, f'Expected {A} actions but {len(flat)} were parsed to env.step.'
        else:
            flat = []
            for b, a in zip(self.spec.buildings, actions):
                a = list(a)
                assert len(a) == len(b.active_actions), f'Expected {len(b.active_actions)} for {b.name} but {len(a)} actions were provided.'
                flat += [float(v) for v in a]
        self._upload(lambda host: host.__setitem__((0, slice(0, A)), np.asarray(flat, dtype=np.float32)))
        return self._act, True

    def _upload(self, fill):
        """Fill the next pinned staging buffer on the host and enqueue its H2D copy into `self._act`."""
        i = self._stage_i
        self._stage_i = (i + 1) % self._stage_n
        if self._stage_used[i]:
            self._stage_event[i].synchronize()          # the copy that last read this buffer has finished
        fill(self._stage_host[i])
        self._act.copy_(self._stage_pinned[i], non_blocking=True)
        self._stage_event[i].record(torch.cuda.current_stream(self.device))
        self._stage_used[i] = True

    def step(self, actions):
        return self._advance(actions, True)

    def _advance(self, actions, write_obs: bool):
        """One time step; `write_obs=False` skips the [E, L] observation slab (shared-row host path)."""
        if self.terminated:
            raise RuntimeError('step() called after the episode terminated; call reset().')
        with torch.cuda.device(self.device):
            a, ref_shaped = self._parse_actions(actions)
            stream = self._stream()
            fused = self._reward_id >= 0
            self._h.step(a.data_ptr(), self._obs.data_ptr() if write_obs else None, self._reward.data_ptr() if fused else None,
                         self._district.data_ptr(), None if self._trace is None else self._trace.data_ptr(), stream)
            self._obs_current = write_obs
            if not fused:
                self._python_reward()
            if self._track_kpis and not self._kpi_fused:
                self._h.kpi_accumulate(self._trace.data_ptr(), self._district.data_ptr(), stream)
            if self._record:
                self._hist_dyn[self.time_step].copy_(self._trace[self._history_env])
                self._hist_district[self.time_step].copy_(self._district[self._history_env])
            self.time_step += 1
            if self._track:
                self._accumulate_rewards()
        terminated = self.terminated
        if terminated and self._track:
            self._finish_episode_rewards()
        if terminated and self.auto_reset:
            final = self._obs.clone() if write_obs else None
            rew = self._reward.clone()
            self.reset()
            return self._obs, rew, True, False, {'final_observation': final}
        if ref_shaped:
            rew = self._reward[0].tolist()
            return self._shape_obs(self._obs), rew, terminated, False, self.get_info()
        return self._obs, self._reward, terminated, False, self.get_info()

    @property
    def shared_observation_row(self) -> bool:
        """True when every env's observation row after a step is the same row: reference-parity (stale) observations and one
        episode window for all envs - the host paths then move ONE row instead of E identical ones."""
        return self.stale_observations and getattr(self, '_uniform_start', True)

    def step_host(self, actions: np.ndarray, full_observations: Optional[bool] = None) -> Tuple[np.ndarray, np.ndarray, bool]:
        """End-to-end host path: host ndarray actions in, host ndarrays out (pinned staging, one H2D + one D2H copy, one sync).

        With `shared_observation_row` (and `full_observations` not forced) the kernel skips the [E, L] observation slab, the
        D2H copy is rewards + ONE observation row, and the returned observations are a read-only broadcast view `[E, L]` of it.
        The returned arrays are views of pinned buffers that the next call overwrites."""
        shared = self.shared_observation_row if full_observations is None else not full_observations
        if shared and not self.shared_observation_row:
            raise ValueError('full_observations=False needs stale_observations=True and one episode window for all envs')
        E, L, R = self.num_envs, self._obs_dim, self._reward_dim
        if self._host_fast and not self.terminated:
            # the whole step in ONE native call (cl_step_host): H2D, kernel, shared row, one D2H, sync - no torch ops on the way
            a = actions if type(actions) is np.ndarray else np.asarray(actions, dtype=np.float32)
            if a.dtype == np.float32 and a.flags.c_contiguous and a.size == self._act.numel():
                # copied to the device straight from the caller's memory: an asynchronous DMA transfer when it is page-locked
                # (`pinned_actions()`), staged by the driver when it is pageable
                src = a.__array_interface__['data'][0]
            else:
                a = np.asarray(a, dtype=np.float32).reshape(E, -1)
                if a.shape[1] != self.spec.action_dim:
                    raise ValueError(f'actions must have shape [{E}, {self.spec.action_dim}]')
                np.copyto(self._stage_host[0], a)
                src = self._stage_pinned[0].data_ptr()
            p = self._host_ptrs
            if torch.cuda.current_device() != self.device.index:
                with torch.cuda.device(self.device):
                    return self.step_host(actions, full_observations)
            st = torch.cuda.current_stream(self.device).cuda_stream
            if shared:
                self._h.step_host(src, p['act'], None, p['reward'], p['district'], p['row'], p['reward'], p['reward_host'], 4 * (E * R + L),
                                  self._host_in_place, st)
                if self._host_in_place & 2:
                    self._reward_current = False            # the device reward buffer was bypassed (results went to the host range)
            else:
                self._h.step_host(src, p['act'], p['obs'], p['reward'], p['district'], None, p['obs'], p['obs_host'], 4 * E * (L + R), 0, st)
            self._obs_current = not shared
            self._hist_valid = False
            self.time_step += 1
            terminated = self.terminated
            if shared:
                return self._row_view, self._reward_host, terminated
            return self._obs_host, self._reward_host, terminated
        _, _, terminated, _, _ = self._advance(np.asarray(actions, dtype=np.float32), not shared)
        with torch.cuda.device(self.device):
            if shared:
                self._h.obs_rows(self.time_step, 1, self._row.data_ptr(), self._stream())
                self._out_pinned[E * L:].copy_(self._out[E * L:], non_blocking=True)
            else:
                self._out_pinned[:E * (L + R)].copy_(self._out[:E * (L + R)], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        if shared:
            return np.broadcast_to(self._row_host, (E, L)), self._reward_host, terminated
        return self._obs_host, self._reward_host, terminated

    def pinned_actions(self, n: int = 1) -> List[np.ndarray]:
        """`n` page-locked `[E, A]` float32 arrays owned by the env (kept alive with it): actions written into one of them reach the device
        by DMA straight from that memory when passed to `step_host` (any page-locked array does; pageable ones are staged by the driver)."""
        out = []
        for _ in range(int(n)):
            t = torch.zeros((self.num_envs, max(self.spec.action_dim, 1)), dtype=torch.float32).pin_memory()
            self._pinned_action_buffers.append(t)
            out.append(t.numpy())
        return out

    def rollout_host(self, actions: np.ndarray, full_observations: Optional[bool] = None):
        """K steps from a host block of actions `[K, E, A]`: one H2D copy, one `cl_rollout` launch, one D2H copy, one sync.
        Returns `(observations, rewards [K, E, R], terminated)` as views of pinned host buffers (overwritten by the next call of the
        same K); observations are `[K, L]` - one row per step - with `shared_observation_row`, else `[K, E, L]`."""
        if self._reward_id < 0:
            raise NotImplementedError('rollout_host needs a built-in (fused) reward function')
        shared = self.shared_observation_row if full_observations is None else not full_observations
        if shared and not self.shared_observation_row:
            raise ValueError('full_observations=False needs stale_observations=True and one episode window for all envs')
        actions = np.asarray(actions, dtype=np.float32)
        K = actions.shape[0]
        E, A, L, R = self.num_envs, self.spec.action_dim, self._obs_dim, self._reward_dim
        actions = actions.reshape(K, E, A)
        if self.time_step + K > self.time_steps - 1:
            raise RuntimeError('rollout_host: the block runs past the end of the episode')
        with torch.cuda.device(self.device):
            key = (K, shared)
            if self._roll is None or self._roll['key'] != key:
                n_obs = K * L if shared else K * E * L
                out = torch.empty(K * E * R + n_obs, dtype=torch.float32, device=self.device)
                out_pinned = torch.empty(K * E * R + n_obs, dtype=torch.float32).pin_memory()
                self._roll = {'key': key, 'act': torch.empty((K, E, A), dtype=torch.float32, device=self.device),
                              'act_pinned': torch.empty((K, E, A), dtype=torch.float32).pin_memory(), 'out': out, 'out_pinned': out_pinned}
            r = self._roll
            np.copyto(r['act_pinned'].numpy(), actions)
            r['act'].copy_(r['act_pinned'], non_blocking=True)
            rew = r['out'][:K * E * R]
            obs = r['out'][K * E * R:]
            stream = self._stream()
            t_first = self.time_step + 1
            self._h.rollout(K, r['act'].data_ptr(), None if shared else obs.data_ptr(), rew.data_ptr(), self._district.data_ptr() if K == 1 else None, stream)
            if shared:
                self._h.obs_rows(t_first, K, obs.data_ptr(), stream)
            r['out_pinned'].copy_(r['out'], non_blocking=True)
            torch.cuda.current_stream(self.device).synchronize()
        self.time_step += K
        self._hist_valid = False
        self._kpi_valid = self._kpi_valid and self._kpi_fused
        self._obs_current = False
        if not shared:
            self._obs.copy_(obs.view(K, E, L)[-1])
            self._obs_current = True
        host = r['out_pinned'].numpy()
        return host[K * E * R:].reshape((K, L) if shared else (K, E, L)), host[:K * E * R].reshape(K, E, R), self.terminated

    def rollout(self, actions: torch.Tensor, obs: Optional[torch.Tensor] = None, reward: Optional[torch.Tensor] = None,
                district: Optional[torch.Tensor] = None):
        """K steps with pre-resident actions `[K, E, A]` (device). Fills `obs [K, E, L]`, `reward [K, E, R]`, `district [K, E, 3]`."""
        if self._reward_id < 0:
            raise NotImplementedError('rollout needs a built-in (fused) reward function')
        K = actions.shape[0]
        assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
        with torch.cuda.device(self.device):
            self._h.rollout(K, actions.data_ptr(), None if obs is None else obs.data_ptr(), None if reward is None else reward.data_ptr(),
                            None if district is None else district.data_ptr(), self._stream())
        self.time_step += K
        self._hist_valid = False          # rollouts do not produce the per-unit trace evaluate() needs
        self._kpi_valid = self._kpi_valid and self._kpi_fused       # fused accumulators run inside every launch
        if obs is not None:
            self._obs.copy_(obs[K - 1])
        self._obs_current = obs is not None
        return obs, reward, self.terminated

    # ---------------------------------------------------------------------------------------------
    def _python_reward(self):
        obs = self._reward_observations()
        r = self.reward_function.calculate(obs)
        flat = []
        for v in r:
            v = v[0] if isinstance(v, (list, tuple)) else v     # MultiBuildingRewardFunction returns 1-element lists
            v = v if isinstance(v, torch.Tensor) else torch.as_tensor(v, dtype=torch.float32, device=self.device)
            flat.append(v.to(self.device, torch.float32).reshape(-1).expand(self.num_envs))
        self._reward.copy_(torch.stack(flat, dim=1))

    def _accumulate_rewards(self):
        r = self._reward
        if self._rsum is None:
            self._rsum, self._rmin, self._rmax, self._rcount = r.clone(), r.clone(), r.clone(), 1
        else:
            self._rsum += r
            torch.minimum(self._rmin, r, out=self._rmin)
            torch.maximum(self._rmax, r, out=self._rmax)
            self._rcount += 1

    def _finish_episode_rewards(self):
        # citylearn/citylearn.py:1034-1040 (float32 min / max / sum / mean over the episode's steps)
        def shape(x):
            return x[0].tolist() if self.num_envs == 1 else x.clone()
        self._episode_rewards.append({'min': shape(self._rmin), 'max': shape(self._rmax), 'sum': shape(self._rsum),
                                      'mean': shape(self._rsum / self._rcount)})

    def evaluate(self, control_condition=None, baseline_condition=None, comfort_band: float = None):
        """KPI table of the recorded env (`citylearn/citylearn.py:1136-1323`): ratios of control vs baseline cost functions per building
        and for the district, as a DataFrame with columns cost_function / value / name / level.  Needs `record_history=True`
        (the default for num_envs == 1) and steps taken through `step()`."""
        from .evaluate import History, evaluate
        if not self._record or not getattr(self, '_hist_valid', False):
            raise RuntimeError('evaluate() needs record_history=True and an episode advanced with step() (not rollout())')
        k = self.time_step
        h = History(self._hist_dyn[:k].cpu().numpy(), self._hist_district[:k].cpu().numpy(), int(self._start_dev[self._history_env].item()),
                    self._outage)
        return evaluate(self.spec, h, control_condition, baseline_condition, comfort_band)

    def evaluate_batched(self) -> Dict[str, Dict[str, np.ndarray]]:
        """The action-dependent KPI ratios of EVERY env (`track_kpis=True`): `{'district': {name: [E]}, 'building': {name: [E, B]}}`,
        control vs the `_without_storage` baseline, from accumulators kept on the device (`cl_kpi_*`) - no per-step history."""
        from .evaluate import evaluate_batched
        if not self._track_kpis or not self._kpi_valid:
            raise RuntimeError('evaluate_batched() needs track_kpis=True (and, for building-tiled districts, an episode advanced with step())')
        E, B = self.num_envs, self.spec.n_buildings
        with torch.cuda.device(self.device):
            unit = torch.empty((E, B, 8), dtype=torch.float64, device=self.device)
            envacc = torch.empty((E, 2, 15), dtype=torch.float64, device=self.device)
            self._h.kpi_read(unit.data_ptr(), envacc.data_ptr(), self._stream())
        return evaluate_batched(self.spec, unit.cpu().numpy(), envacc.cpu().numpy())

    # ---------------------------------------------------------------------------------------------
    def state_dict(self) -> Dict[str, Any]:
        """Checkpoint of the mutable simulation state and of the episode it belongs to."""
        with torch.cuda.device(self.device):
            n = self._h.state_size()
            buf = torch.empty(n, dtype=torch.uint8, device=self.device)
            self._h.get_state(buf.data_ptr(), self._stream())
        _ = self.observations                  # refreshes self._obs after shared-row host steps
        obs = self._obs.clone()
        tr = self.episode_tracker
        return {'state': buf, 'time_step': self.time_step, 'episode': tr.episode,
                'episode_window': (tr.episode_start_time_step, tr.episode_end_time_step), 'uniform_start': self._uniform_start,
                'episode_start': self._start_dev.clone(), 'outage': np.array(self._outage, copy=True), 'obs': obs}

    def load_state_dict(self, sd: Mapping[str, Any]):
        """Resume a checkpoint: the episode window(s), the outage signal and the device state are all restored, so loading into a fresh
        env or after a reset() that picked another window continues on the checkpoint's time-series rows.  Per-step history and KPI
        accumulators are not part of a checkpoint: `evaluate()` / `evaluate_batched()` are unavailable until the next reset()."""
        start, end = (int(v) for v in sd['episode_window'])
        T = end - start + 1
        starts = torch.as_tensor(sd['episode_start']).to(device=self.device, dtype=torch.int32).contiguous()
        if starts.shape != (self.num_envs,):
            raise ValueError(f'checkpoint holds {tuple(starts.shape)} episode starts, this env has {self.num_envs} envs')
        if int(starts.max()) + T > self.spec.table.shape[0] or int(starts.min()) < 0:
            raise ValueError('checkpoint episode window lies outside this dataset')
        if not 0 <= int(sd['time_step']) <= T - 1:
            raise ValueError('checkpoint time step outside its episode window')
        tr = self.episode_tracker
        tr.episode, tr.episode_start_time_step, tr.episode_end_time_step = int(sd['episode']), start, end
        self._outage = np.array(sd['outage'], copy=True)
        with torch.cuda.device(self.device):
            stream = self._stream()
            self._h.set_outage(self._outage if any(b.simulate_power_outage for b in self.spec.buildings) else None, stream)
            self._start_dev = starts
            self._uniform_start = bool(sd.get('uniform_start', True))
            if self._uniform_start:
                self._h.reset(None, start, T, None, stream)
            else:
                self._h.reset(self._start_dev.data_ptr(), 0, T, None, stream)
            self._h.set_state(sd['state'].data_ptr(), int(sd['time_step']), stream)
            self.time_step = int(sd['time_step'])
            self._obs.copy_(sd['obs'])
        self._obs_current = True
        self._hist_valid = False
        self._kpi_valid = False
        self._rsum = self._rmin = self._rmax = None

    @property
    def trace(self) -> Optional[torch.Tensor]:
        """`[E, B, CL_NDYN]` per-unit values of the last step (only with `debug_trace=True` or a Python reward function)."""
        return self._trace

    @property
    def district(self) -> torch.Tensor:
        """`[E, 3]` district net electricity consumption, cost and emission of the last step (citylearn.py:1908-1918)."""
        return self._district

    @property
    def gpu_launches(self) -> int:
        return self._h.launch_count()
