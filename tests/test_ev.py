"""Electric vehicles, chargers, washing machines: loader and oracle against traces of the UNMODIFIED reference
(tests/golden/ev/*.npz, recorded by oracle/make_golden.py with NumPy's global generator seeded - the reference draws the SOC drift of
away vehicles from it), GPU parity through the C ABI."""
import json

import numpy as np
import pytest

from citylearn_b200 import schema as S
from citylearn_b200.data import DataSet
from citylearn_oracle import OracleEnv
from helpers import GOLDEN

CASES = sorted(p.stem for p in (GOLDEN / 'ev').glob('*.npz'))


def load(case):
    z = np.load(GOLDEN / 'ev' / f'{case}.npz')
    cfg = json.loads(bytes(z['config']).decode())
    meta = json.loads(bytes(z['meta']).decode())
    src = DataSet.get_source(cfg['dataset'])
    sch = src.schema()
    if cfg.get('reward') is not None:
        sch['reward_function'] = {'type': cfg['reward']['type'], 'attributes': cfg['reward'].get('attributes', {})}
    spec = S.load(sch, data_source=src, ev_random_seed=cfg['np_seed'], **(cfg.get('overrides') or {}))
    return z, cfg, meta, spec


@pytest.mark.parametrize('case', CASES)
def test_loader_reproduces_reference_names_and_spaces(case):
    z, cfg, meta, spec = load(case)
    assert [list(b.active_observations) for b in spec.buildings] == meta['observation_names']
    assert [list(b.active_actions) for b in spec.buildings] == meta['action_names']
    for b, lo, hi, alo, ahi in zip(spec.buildings, meta['observation_low'], meta['observation_high'], meta['action_low'], meta['action_high']):
        assert np.array_equal(b.observation_low, np.float32(lo)) and np.array_equal(b.observation_high, np.float32(hi))
        assert np.array_equal(b.action_low, np.float32(alo)) and np.array_equal(b.action_high, np.float32(ahi))
    assert [e.name for e in spec.evs] == cfg['vehicles'] and [c.charger_id for c in spec.ev['chargers']] == cfg['chargers']


@pytest.mark.parametrize('case', CASES)
def test_oracle_matches_reference_bit_for_bit(case):
    """Observations, rewards (incl. Electric_Vehicles_Reward_Function), district sums, every vehicle's SOC entry, every charger's and
    washing machine's consumption and the charged energy - exact at every recorded step (exact-zero actions included)."""
    z, cfg, meta, spec = load(case)
    env = OracleEnv(spec, 1, libm_pow=True)
    assert np.array_equal(env.reset()[0].astype('float32'), z['reset_obs'])
    for k in range(len(z['actions'])):
        obs, rew, dist, dyn = env.step(z['actions'][k][None])
        info = env.last_ev
        assert np.array_equal(obs[0], z['obs'][k]), k
        assert np.array_equal(rew[0], z['reward'][k]), k
        assert np.array_equal(dist[0], z['district'][k]), k
        assert np.array_equal(info['ch_ec'][0], z['charger_ec'][k]) and np.array_equal(info['past'][0], z['charger_kwh'][k]), k
        assert np.array_equal(info['wm_ec'][0], z['wm_ec'][k]), k
        assert np.array_equal(env.ev_soc_prev[0].astype('float32'), z['ev_soc'][k]), k     # soc[k], one `next_time_step` later


def test_schedule_is_action_independent_and_seeded():
    """Two loads with the same `ev_random_seed` agree, another seed changes only the away-drift factors."""
    a = S.load('citylearn_challenge_2022_phase_all_plus_evs', ev_random_seed=3).ev['schedule']
    b = S.load('citylearn_challenge_2022_phase_all_plus_evs', ev_random_seed=3).ev['schedule']
    c = S.load('citylearn_challenge_2022_phase_all_plus_evs', ev_random_seed=4).ev['schedule']
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True)
    assert np.array_equal(a['assoc'], c['assoc'], equal_nan=True) and not np.array_equal(a['drift'], c['drift'], equal_nan=True)
    d = a['drift'][np.isfinite(a['drift'])]
    assert d.size > 0 and d.min() >= 0.6 and d.max() <= 1.4
