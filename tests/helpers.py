"""Shared test helpers: golden fixtures, spec construction from a fixture's config, comparison utilities."""
import functools
import json
from pathlib import Path

import numpy as np

from citylearn_b200 import schema as S
from citylearn_b200.data import DataSet

GOLDEN = Path(__file__).resolve().parent / 'golden'

# golden trace name -> cl_dyn name
TRACE_TO_DYN = {
    'electrical_storage_soc': 'electrical_storage_soc',
    'electrical_storage_energy_balance': 'electrical_storage_energy_balance',
    'electrical_storage_electricity_consumption': 'electrical_storage_electricity_consumption',
    'electrical_storage_degraded_capacity': 'electrical_storage_degraded_capacity',
    'non_shiftable_load_electricity_consumption': 'non_shiftable_load_electricity_consumption',
    'cooling_electricity_consumption': 'cooling_electricity_consumption',
    'heating_electricity_consumption': 'heating_electricity_consumption',
    'dhw_electricity_consumption': 'dhw_electricity_consumption',
    'cooling_storage_soc': 'cooling_storage_soc', 'heating_storage_soc': 'heating_storage_soc', 'dhw_storage_soc': 'dhw_storage_soc',
    'cooling_storage_energy_balance': 'cooling_storage_energy_balance', 'heating_storage_energy_balance': 'heating_storage_energy_balance',
    'dhw_storage_energy_balance': 'dhw_storage_energy_balance',
    'net_electricity_consumption': 'net_electricity_consumption',
    'net_electricity_consumption_cost': 'net_electricity_consumption_cost',
    'net_electricity_consumption_emission': 'net_electricity_consumption_emission',
    'indoor_dry_bulb_temperature': 'indoor_dry_bulb_temperature',
}


def golden_cases():
    return sorted(p.stem for p in GOLDEN.glob('*.npz'))


def load_golden(name):
    z = np.load(GOLDEN / f'{name}.npz')
    cfg = json.loads(bytes(z['config']).decode())
    meta = json.loads(bytes(z['meta']).decode())
    return z, cfg, meta


def schema_for(cfg):
    """(schema dict, data source, overrides) reproducing the fixture's environment from the bundled packs."""
    if cfg['dataset'].startswith('synthetic_wide_'):
        from citylearn_b200.synthetic import SyntheticWideSource
        src = SyntheticWideSource(int(cfg['dataset'].rsplit('_', 1)[1]))
    elif cfg['dataset'] == 'synthetic_dual_mode':
        from citylearn_b200.synthetic import SyntheticDualModeSource
        src = SyntheticDualModeSource()
    elif cfg['dataset'] == 'synthetic_heating':
        from citylearn_b200.synthetic import SyntheticHeatingSource
        src = SyntheticHeatingSource()
    else:
        src = DataSet.get_source(cfg['dataset'])
    sch = src.schema()
    if cfg.get('reward') is not None:
        sch['reward_function'] = {'type': cfg['reward']['type'], 'attributes': cfg['reward'].get('attributes', {})}
    return sch, src, dict(cfg.get('overrides') or {})


def spec_for(cfg):
    sch, src, ov = schema_for(cfg)
    return S.load(sch, data_source=src, **ov)


def actions_of(z):
    a = z['actions']
    return a if a.ndim == 3 else a[None]


def max_abs_diff(a, b):
    a = np.asarray(a, dtype='float64')
    b = np.asarray(b, dtype='float64')
    assert np.array_equal(np.isnan(a), np.isnan(b)), 'NaN pattern differs'
    m = ~np.isnan(a)
    return float(np.max(np.abs(a[m] - b[m]))) if m.any() else 0.0


def within_scaled_tolerance(x, ref, scale, rtol=1e-5):
    """|x - ref| <= rtol * max(|ref|, scale)  (SURVEY.md §8c); NaNs must coincide."""
    x = np.asarray(x, dtype='float64')
    ref = np.asarray(ref, dtype='float64')
    scale = np.broadcast_to(np.asarray(scale, dtype='float64'), ref.shape)
    if not np.array_equal(np.isnan(x), np.isnan(ref)):
        return False, float('inf')
    m = ~np.isnan(ref)
    bound = rtol * np.maximum(np.abs(ref[m]), scale[m])
    err = np.abs(x[m] - ref[m])
    worst = float(np.max(err / bound)) if m.any() else 0.0
    return bool(np.all(err <= bound)), worst


def observation_scales(spec, entries):
    """per-observation scale = high - low of its space (at least 1)."""
    out = []
    for bi, name in entries:
        b = spec.buildings[bi]
        k = b.active_observations.index(name)
        s = float(b.observation_high[k]) - float(b.observation_low[k])
        out.append(max(s, 1.0) if np.isfinite(s) else 1.0)
    return np.array(out)


@functools.lru_cache(maxsize=None)
def oracle_run(case):
    """One oracle episode over a single-episode fixture's actions (shared by test_oracle_golden and test_evaluate: the year-long
    cases take ~40 s each).  `libm_pow=True`: the reference's `efficiency ** 0.5` (see OracleEnv)."""
    from citylearn_oracle import OracleEnv
    z, cfg, _ = load_golden(case)
    assert cfg['episodes'] == 1
    spec = spec_for(cfg)
    env = OracleEnv(spec, 1, libm_pow=True)
    tracker = S.EpisodeTracker(spec.simulation_start_time_step, spec.simulation_end_time_step)
    ets = spec.episode_time_steps if spec.episode_time_steps is not None else tracker.simulation_time_steps
    tracker.next_episode(ets, spec.rolling_episode_split, spec.random_episode_split, spec.random_seed)
    reset_obs = env.reset(tracker.episode_start_time_step, tracker.episode_time_steps)[0].astype('float32')
    acts = actions_of(z)[0]
    obs, rew, dist, dyn = [], [], [], []
    for k in range(len(acts)):
        o, r, d, y = env.step(acts[k][None])
        obs.append(o[0].astype('float32')); rew.append(r[0].copy()); dist.append(d[0].copy()); dyn.append(y[0].copy())
    return {'spec': spec, 'window': [tracker.episode_start_time_step, tracker.episode_end_time_step], 'reset_obs': reset_obs,
            'obs': np.stack(obs), 'reward': np.stack(rew), 'district': np.stack(dist), 'dyn': np.stack(dyn),
            'start': int(env.start[0]), 'outage': env.outage.copy()}
