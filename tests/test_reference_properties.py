"""Invariants the reference's own test-suite pins for this path, restated against the oracle / loader.

reference tests mirrored:
  tests/unit/test_alignment.py:47-60 ......... net[t] equals the sum of its component series (1e-4)
  tests/unit/test_battery.py:135-207 ........ SOC rises on charge, DoD floor, capacity ceiling, degradation formula
  tests/unit/test_pv.py:22-32 ............... PV.get_generation exact values
  tests/test_series_integrity.py:41-52 ...... stale-zero observations after a step; T-1 steps per episode
  tests/unit/test_subhour_scaling.py:65-82 .. time_step_ratio semantics (golden case c1_subhour)
"""
import numpy as np
import pytest

from citylearn_b200 import schema as S
from citylearn_b200.schema import DYN, P
from citylearn_oracle import OracleEnv


@pytest.fixture(scope='module')
def spec():
    return S.load('citylearn_challenge_2022_phase_1')


def test_net_equals_sum_of_components(spec):
    env = OracleEnv(spec, 4)
    env.reset()
    rng = np.random.RandomState(1)
    for k in range(30):
        _, _, _, dyn = env.step(rng.uniform(-1, 1, size=(4, spec.action_dim)).astype('float32'))
        comp = (dyn[..., DYN['cooling_electricity_consumption']] + dyn[..., DYN['heating_electricity_consumption']]
                + dyn[..., DYN['dhw_electricity_consumption']] + dyn[..., DYN['non_shiftable_load_electricity_consumption']]
                + dyn[..., DYN['electrical_storage_electricity_consumption']] + env.solar64(k))
        assert np.max(np.abs(comp - dyn[..., DYN['net_electricity_consumption']])) < 1e-4


def test_battery_invariants(spec):
    E = 3
    env = OracleEnv(spec, E)
    env.reset()
    cap = spec.params[:, P['BAT_CAPACITY']]
    prev_soc = np.broadcast_to(spec.params[:, P['BAT_INITIAL_SOC']], (E, spec.n_buildings)).copy()
    prev_cap = np.broadcast_to(cap, (E, spec.n_buildings)).copy()
    acts = [np.full((E, spec.action_dim), v, dtype='float32') for v in (0.5, 0.5, 1.0, 1.0, -0.3, -1.0, -1.0, -1.0, 0.2)]
    for a in acts:
        _, _, _, dyn = env.step(a)
        soc = dyn[..., DYN['electrical_storage_soc']]
        eb = dyn[..., DYN['electrical_storage_energy_balance']]
        deg = dyn[..., DYN['electrical_storage_degraded_capacity']]
        assert np.all(soc >= -1e-7) and np.all(soc <= 1.0 + 1e-6)                       # capacity ceiling / floor
        if a[0, 0] > 0:
            assert np.all(soc >= prev_soc - 1e-7) and np.all(eb >= 0)                   # SOC rises on charge
        else:
            assert np.all(soc <= prev_soc + 1e-7) and np.all(eb <= 0)
        assert np.all(soc >= (1.0 - spec.params[:, P['BAT_DOD']]) - 1e-6)               # depth-of-discharge floor
        # degradation: capacity_loss_coefficient * capacity * |energy_balance| / (2 * degraded_capacity)
        expect = prev_cap - spec.params[:, P['BAT_CLC']] * cap * np.abs(eb) / (2 * np.maximum(prev_cap, 1e-6))
        np.testing.assert_allclose(deg, np.maximum(expect, 0), rtol=1e-6)
        assert np.all(deg <= prev_cap + 1e-12)
        prev_soc, prev_cap = soc, deg


def test_pv_generation_exact():
    # reference tests/unit/test_pv.py:22-32: generation = nominal_power * inverter_ac_power_per_kw / 1000
    b = S.load('citylearn_challenge_2022_phase_1').buildings[0]
    series = np.array([0.0, 250.0, 1000.0], dtype='float32')
    np.testing.assert_array_equal(S.pv_generation(b, series), b.devices['pv']['nominal_power'] * series.astype('float64') / 1000.0)


def test_episode_has_T_minus_1_steps_and_stale_observations(spec):
    short = S.load('citylearn_challenge_2022_phase_1', simulation_end_time_step=23)
    env = OracleEnv(short, 1)
    env.reset()
    names = [n for _, n in env.entries]
    i_soc, i_net = names.index('electrical_storage_soc'), names.index('net_electricity_consumption')
    steps = 0
    while env.t < env.T - 1:                      # terminated <=> time_step == time_steps - 1 (citylearn.py:372-376)
        obs, _, _, dyn = env.step(np.full((1, short.action_dim), 0.7, dtype='float32'))
        steps += 1
        assert obs[0, i_soc] == 0.0 and obs[0, i_net] == 0.0
        assert dyn[0, 0, DYN['electrical_storage_soc']] > 0.0          # ... although the battery did charge
    assert steps == 23


def test_first_step_multicounting(spec):
    """t = 0: thermal devices and the non-shiftable load are counted three times, the battery twice."""
    env = OracleEnv(spec, 1)
    env.reset()
    a = np.full((1, spec.action_dim), 0.1, dtype='float32')
    _, _, _, dyn = env.step(a)
    nsl0 = env.col('C_NSL', 0)[0]
    np.testing.assert_allclose(dyn[0, :, DYN['non_shiftable_load_electricity_consumption']], 3 * nsl0, rtol=1e-6)
    np.testing.assert_allclose(dyn[0, :, DYN['electrical_storage_electricity_consumption']],
                               2 * dyn[0, :, DYN['electrical_storage_energy_balance']], rtol=1e-6)
