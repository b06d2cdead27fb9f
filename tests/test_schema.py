"""Host loader vs metadata recorded from the unmodified reference (tests/golden/*.npz 'meta')."""
import numpy as np
import pytest

from citylearn_b200 import schema as S
from helpers import golden_cases, load_golden, spec_for


@pytest.mark.parametrize('case', golden_cases())
def test_names_spaces_and_devices_match_reference(case):
    z, cfg, meta = load_golden(case)
    spec = spec_for(cfg)
    entries, desc = S.observation_layout(spec)
    assert [n for _, n in entries] == [n for row in meta['observation_names'] for n in row]
    assert spec.central_agent == meta['central_agent']
    assert spec.shared_observations == meta['shared_observations']
    for b, mb in zip(spec.buildings, meta['buildings']):
        assert b.name == mb['name']
        assert b.active_observations == mb['active_observations']
        assert b.active_actions == mb['active_actions']
        np.testing.assert_allclose(b.observation_low, mb['observation_low'], rtol=1e-6, equal_nan=True)
        np.testing.assert_allclose(b.observation_high, mb['observation_high'], rtol=1e-6, equal_nan=True)
        np.testing.assert_allclose(b.action_low, mb['action_low'], rtol=1e-6)
        np.testing.assert_allclose(b.action_high, mb['action_high'], rtol=1e-6)
        bat, mbat = b.devices['electrical_storage'], mb['electrical_storage']
        # md5-seeded stochastic defaults (citylearn/citylearn.py:2364-2378, energy_model.py:977-1003)
        if not bat.get('absent'):   # an absent battery is a zero-sized one whose defaults the reference draws from the global `random`
            np.testing.assert_allclose(bat['power_efficiency_curve'], mbat['power_efficiency_curve'], rtol=1e-13)
            np.testing.assert_allclose(bat['capacity_power_curve'], mbat['capacity_power_curve'], rtol=1e-13)
            for k in ('capacity', 'nominal_power', 'depth_of_discharge', 'capacity_loss_coefficient', 'initial_soc'):
                assert bat[k] == pytest.approx(mbat[k], rel=1e-13)
        else:
            assert mbat['capacity'] == 0.0 and mbat['nominal_power'] == 0.0
        for dn in ('cooling_storage', 'heating_storage', 'dhw_storage', 'cooling_device', 'heating_device', 'dhw_device'):
            if b.devices[dn].get('absent'):
                continue          # absent devices get unreproducible random parameters in the reference
            for k, v in b.devices[dn].items():
                if k in mb[dn] and isinstance(v, float):
                    assert v == pytest.approx(mb[dn][k], rel=1e-13), (dn, k)


def test_device_seed_known_answer():
    # 2022 Building_1 battery, schema seed 2022 -> 135823784 -> u = 0.5803480936676708
    seed = S.device_random_seed('Building_1', 'citylearn.citylearn.Building', 'electrical_storage', 'citylearn.energy_model.Battery', 2022)
    assert seed == 135823784
    assert np.random.RandomState(seed).uniform() == pytest.approx(0.5803480936676708, rel=1e-15)


def test_episode_windows_match_reference():
    z, cfg, meta = load_golden('c1_episodes')
    spec = spec_for(cfg)
    tr = S.EpisodeTracker(spec.simulation_start_time_step, spec.simulation_end_time_step)
    for ep in range(cfg['episodes']):
        tr.next_episode(spec.episode_time_steps, spec.rolling_episode_split, spec.random_episode_split, spec.random_seed)
        assert [tr.episode_start_time_step, tr.episode_end_time_step] == z['episode_window'][ep].tolist()


def test_unsupported_features_fail_loudly():
    from citylearn_b200.data import DataSet
    src = DataSet.get_source('citylearn_challenge_2022_phase_1')
    sch = src.schema()
    sch['buildings']['Building_1']['chargers'] = {'c1': {}}
    with pytest.raises(ValueError, match='charger_simulation'):
        S.load(sch, data_source=src)
    # chargers are supported (citylearn_b200/ev.py); load-time noise from NumPy's global generator is not
    ev_src = DataSet.get_source('citylearn_challenge_2022_phase_all_plus_evs')
    sch = ev_src.schema()
    b = next(n for n, v in sch['buildings'].items() if v.get('chargers'))
    next(iter(sch['buildings'][b]['chargers'].values()))['noise_std'] = 0.1
    with pytest.raises(S.UnsupportedSchemaError):
        S.load(sch, data_source=ev_src)
    with pytest.raises(S.UnknownSchemaError):
        S.load('no_such_dataset')


def test_table_layout():
    spec = S.load('citylearn_challenge_2022_phase_all')
    assert spec.table.dtype == np.float32 and spec.table.shape[0] == 8760
    assert spec.params.shape == (17, S.NPARAM) and spec.iparams.shape == (17, S.NIPARAM)
    assert spec.action_dim == 17
    # shared weather / pricing series are stored once
    assert spec.columns[(0, 'outdoor_dry_bulb_temperature')] == spec.columns[(16, 'outdoor_dry_bulb_temperature')]


def test_synthetic_wide_district_directory_round_trip(tmp_path):
    """The synthetic C4 district written as a real schema directory loads to the same tables as the in-memory source."""
    from citylearn_b200.synthetic import SyntheticWideSource
    src = SyntheticWideSource(6)
    root = src.write_directory(tmp_path / 'wide6')
    a = S.load(src.schema(), data_source=src)
    b = S.load(str(root / 'schema.json'))
    assert a.table.shape == b.table.shape and np.array_equal(a.table, b.table, equal_nan=True)
    assert np.array_equal(a.params, b.params, equal_nan=True) and np.array_equal(a.iparams, b.iparams)
    # building i replays Building_{(i mod 17)+1} with a scaled load and a 4 / 5 kW PV
    assert [bb.devices['pv']['nominal_power'] for bb in a.buildings] == [4.0, 5.0, 4.0, 5.0, 4.0, 5.0]
    base = S.load('citylearn_challenge_2022_phase_all')
    ratio = a.buildings[3].series['non_shiftable_load'] / np.maximum(base.buildings[3].series['non_shiftable_load'], 1e-9)
    assert np.nanstd(ratio[base.buildings[3].series['non_shiftable_load'] > 0.1]) < 1e-6


def test_loaded_spec_passes_through_load():
    spec = S.load('citylearn_challenge_2022_phase_1')
    assert S.load(spec) is spec
    with pytest.raises(ValueError):
        S.load(spec, central_agent=True)


def test_every_bundled_dataset_loads_and_steps():
    """All datasets shipped as packs resolve by name (the reference would download them, citylearn/data.py:113-189), load through
    the schema loader and advance a few oracle steps with finite results; EV / occupant datasets are not bundled (unsupported)."""
    from citylearn_b200.data import DataSet
    from citylearn_oracle import OracleEnv
    names = DataSet.get_dataset_names()
    assert len(names) >= 18
    for name in names:
        kw = {'buildings': ['Building_1', 'Building_2', 'Building_3']} if name == 'baeda_3dem' else {}     # Building_4: 1 x 50 LSTM
        spec = S.load(name, **kw)
        assert spec.table.dtype == np.float32 and spec.n_buildings >= 1, name
        env = OracleEnv(spec, 2)
        obs = env.reset()
        assert obs.shape == (2, len(S.observation_layout(spec)[0])), name
        rng = np.random.RandomState(0)
        for _ in range(3):
            lo = np.array([v for b in spec.buildings for v in b.action_low]); hi = np.array([v for b in spec.buildings for v in b.action_high])
            a = (lo + rng.uniform(size=(2, spec.action_dim)) * (hi - lo)).astype('float32')
            obs, rew, dist, dyn = env.step(a)
            assert np.isfinite(rew).all() and np.isfinite(dist).all(), name


def test_loader_fuzz_matches_reference_metadata():
    """24 random combinations of constructor overrides (central agent, building subsets by index / name, inactive observations,
    simulation windows, episode splits, shared observations) over six datasets: names, spaces and the first three episode windows
    equal what the unmodified reference reports (tests/golden/meta_fuzz.json, oracle/make_golden.py meta_fuzz)."""
    import json
    from helpers import GOLDEN
    cases = json.load(open(GOLDEN / 'meta_fuzz.json'))['cases']
    assert len(cases) >= 20
    for c in cases:
        spec = S.load(c['dataset'], **c['overrides'])
        tag = (c['dataset'], c['overrides'])
        assert [b.name for b in spec.buildings] == c['building_names'], tag
        assert spec.central_agent == c['central_agent'], tag
        entries, _ = S.observation_layout(spec)
        assert [n for _, n in entries] == [n for row in c['observation_names'] for n in row], tag
        assert [n for b in spec.buildings for n in b.active_actions] == [n for row in c['action_names'] for n in row], tag
        lo = {(bi, n): v for bi, b in enumerate(spec.buildings) for n, v in zip(b.active_observations, b.observation_low)}
        hi = {(bi, n): v for bi, b in enumerate(spec.buildings) for n, v in zip(b.active_observations, b.observation_high)}
        np.testing.assert_allclose([lo[e] for e in entries], [v for row in c['observation_low'] for v in row], rtol=1e-6, err_msg=str(tag))
        np.testing.assert_allclose([hi[e] for e in entries], [v for row in c['observation_high'] for v in row], rtol=1e-6, err_msg=str(tag))
        np.testing.assert_allclose([v for b in spec.buildings for v in b.action_low], [v for row in c['action_low'] for v in row], rtol=1e-6, err_msg=str(tag))
        np.testing.assert_allclose([v for b in spec.buildings for v in b.action_high], [v for row in c['action_high'] for v in row], rtol=1e-6, err_msg=str(tag))
        tr = S.EpisodeTracker(spec.simulation_start_time_step, spec.simulation_end_time_step)
        for w in c['windows']:
            ets = spec.episode_time_steps if spec.episode_time_steps is not None else tr.simulation_time_steps
            tr.next_episode(ets, spec.rolling_episode_split, spec.random_episode_split, spec.random_seed)
            assert [tr.episode_start_time_step, tr.episode_end_time_step] == w, tag
