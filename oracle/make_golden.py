"""Generate golden traces by running the UNMODIFIED reference (`/root/reference`) on CPU.

TEST INFRASTRUCTURE ONLY (never imported by the product).  Runs in the build container, where the
reference tree exists; the GPU box only ever sees the committed `.npz` fixtures in `tests/golden/`.

    python oracle/make_golden.py            # regenerate every fixture
    python oracle/make_golden.py c2_year    # one case

The reference needs two pure-Python packages that are absent from this image (`gymnasium`,
`simplejson`); `oracle/shims/` holds minimal stand-ins.  The dataset cache is
seeded from `/root/reference/data/misc` so that `DataSet()` never touches the network
(`citylearn/citylearn.py:2055-2057`).

Fixture format (one `.npz` per case): `config` (JSON: dataset, overrides, reward), `actions`
[K, sum(A_b)] float32 fed as exact Python floats, `steps` (indices of recorded steps), per-step arrays
`obs` [n, L], `reward` [n, R], `district` [n, 3], per-building internals `trace` [n, B, NTRACE] in
`TRACE_NAMES` order, plus `meta` (JSON: names, spaces, resolved device parameters).
"""
import json
import logging
import os
import shutil
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path('/root/reference')
DATASETS = REF / 'data' / 'datasets'
OUT = HERE.parent / 'tests' / 'golden'

TRACE_NAMES = [
    'electrical_storage_soc', 'electrical_storage_energy_balance', 'electrical_storage_electricity_consumption',
    'electrical_storage_efficiency', 'electrical_storage_degraded_capacity',
    'non_shiftable_load_electricity_consumption', 'cooling_electricity_consumption', 'heating_electricity_consumption',
    'dhw_electricity_consumption', 'cooling_storage_soc', 'heating_storage_soc', 'dhw_storage_soc',
    'cooling_storage_energy_balance', 'heating_storage_energy_balance', 'dhw_storage_energy_balance',
    'net_electricity_consumption', 'net_electricity_consumption_cost', 'net_electricity_consumption_emission',
    'indoor_dry_bulb_temperature', 'cooling_demand', 'heating_demand', 'energy_from_cooling_device',
    'energy_from_dhw_device', 'power_outage',
]


def import_reference():
    sys.path.insert(0, str(HERE / 'shims'))
    sys.path.insert(0, str(REF))
    from platformdirs import user_cache_dir
    d = user_cache_dir(appname='citylearn', appauthor='intelligent-environments-lab', version='v2.4.2')
    os.makedirs(d + '/misc', exist_ok=True)
    for f in ('battery_choices.yaml', 'lbl-tracking_the_sun-res-pv.csv'):
        if not os.path.isfile(d + '/misc/' + f):
            shutil.copy(REF / 'data' / 'misc' / f, d + '/misc/' + f)
    logging.getLogger().setLevel(logging.WARNING)
    from citylearn.citylearn import CityLearnEnv
    return CityLearnEnv


def make_env(CityLearnEnv, dataset, overrides=None, reward=None, schema_hook=None):
    # `EnergySimulation.__init__(..., time_step_ratios=[])` is a mutable default shared by every env of the process and indexed by
    # building position (citylearn/data.py:403,454): a second env in the same process silently reuses the FIRST env's ratios.
    # Clear it so that every fixture is what a fresh process would produce.
    from citylearn.data import EnergySimulation
    for dflt in (EnergySimulation.__init__.__defaults__ or ()):
        if isinstance(dflt, list):
            dflt.clear()
    overrides = dict(overrides or {})
    if dataset.startswith('synthetic_wide_'):
        # first N buildings of the synthetic wide district (BASELINE.json configs[3]), written out as a real schema directory
        import tempfile
        sys.path.insert(0, str(HERE.parent))
        from citylearn_b200.synthetic import SyntheticWideSource
        root = SyntheticWideSource(int(dataset.rsplit('_', 1)[1])).write_directory(tempfile.mkdtemp(prefix='citylearn_wide_'))
    elif dataset == 'synthetic_dual_mode':
        import tempfile
        sys.path.insert(0, str(HERE.parent))
        from citylearn_b200.synthetic import SyntheticDualModeSource
        root = SyntheticDualModeSource().write_directory(tempfile.mkdtemp(prefix='citylearn_dual_'))
    elif dataset == 'synthetic_heating':
        import tempfile
        sys.path.insert(0, str(HERE.parent))
        from citylearn_b200.synthetic import SyntheticHeatingSource
        root = SyntheticHeatingSource().write_directory(tempfile.mkdtemp(prefix='citylearn_heating_'))
    else:
        root = DATASETS / dataset
    schema = json.load(open(root / 'schema.json'))
    schema['root_directory'] = str(root)
    if reward is not None:   # set in the schema dict so that schema attributes do not leak
        schema['reward_function'] = {'type': reward['type'], 'attributes': reward.get('attributes', {})}
    if schema_hook is not None:
        schema_hook(schema)
    return CityLearnEnv(schema, **overrides)


def unit_trace(b, t):
    es = b.electrical_storage
    def at(a):
        return float(a[t])
    return [
        at(es.soc), at(es.energy_balance), at(es.electricity_consumption), float(es.efficiency_history[min(t + 1, len(es.efficiency_history) - 1)]),
        float(es.capacity_history[min(t + 1, len(es.capacity_history) - 1)]),
        at(b.non_shiftable_load_device.electricity_consumption), at(b.cooling_device.electricity_consumption),
        at(b.heating_device.electricity_consumption), at(b.dhw_device.electricity_consumption),
        at(b.cooling_storage.soc), at(b.heating_storage.soc), at(b.dhw_storage.soc),
        at(b.cooling_storage.energy_balance), at(b.heating_storage.energy_balance), at(b.dhw_storage.energy_balance),
        at(b.net_electricity_consumption), at(b.net_electricity_consumption_cost), at(b.net_electricity_consumption_emission),
        at(b.energy_simulation.indoor_dry_bulb_temperature), at(b.energy_simulation.cooling_demand),
        at(b.energy_simulation.heating_demand), at(b.energy_from_cooling_device), at(b.energy_from_dhw_device),
        float(b.power_outage_signal[t]),
    ]


def device_meta(d):
    out = {}
    for k, v in d.get_metadata().items():
        if isinstance(v, np.ndarray):
            out[k] = v.tolist()
        elif isinstance(v, (int, float, str, bool)) or v is None:
            out[k] = v
        elif isinstance(v, np.generic):
            out[k] = v.item()
    out['class'] = type(d).__name__
    return out


def env_meta(env):
    meta = {
        'observation_names': env.observation_names, 'action_names': env.action_names,
        'observation_low': [s.low.tolist() for s in env.observation_space],
        'observation_high': [s.high.tolist() for s in env.observation_space],
        'action_low': [s.low.tolist() for s in env.action_space], 'action_high': [s.high.tolist() for s in env.action_space],
        'central_agent': env.central_agent, 'shared_observations': env.shared_observations,
        'time_steps': env.time_steps, 'seconds_per_time_step': env.seconds_per_time_step, 'random_seed': env.random_seed,
        'buildings': [],
    }
    for b in env.buildings:
        meta['buildings'].append({
            'name': b.name, 'active_observations': b.active_observations, 'active_actions': b.active_actions,
            'observation_low': b.observation_space.low.tolist(), 'observation_high': b.observation_space.high.tolist(),
            'action_low': b.action_space.low.tolist(), 'action_high': b.action_space.high.tolist(),
            **{dn: device_meta(getattr(b, dn)) for dn in ('cooling_device', 'heating_device', 'dhw_device', 'cooling_storage',
                                                            'heating_storage', 'dhw_storage', 'electrical_storage', 'pv')},
            'time_step_ratio': b.time_step_ratio,
        })
    return meta


def flat(list_of_lists):
    return np.array([float(v) for row in list_of_lists for v in (row if isinstance(row, (list, tuple, np.ndarray)) else [row])], dtype='float64')


def run_case(CityLearnEnv, name, dataset, overrides=None, reward=None, steps=None, seed=0, record=None, episodes=1, resets_with_seed=None):
    env = make_env(CityLearnEnv, dataset, overrides, reward)
    meta = env_meta(env)
    lo = np.concatenate([np.asarray(b.action_space.low, dtype='float64') for b in env.buildings])
    hi = np.concatenate([np.asarray(b.action_space.high, dtype='float64') for b in env.buildings])
    sizes = [b.action_space.shape[0] for b in env.buildings]
    rng = np.random.RandomState(seed)
    out = {'obs': [], 'reward': [], 'district': [], 'trace': [], 'steps': [], 'episode': [], 'reset_obs': [], 'episode_window': [],
           'terminated': []}
    all_actions = []
    for ep in range(episodes):
        obs, _ = env.reset()
        out['reset_obs'].append(flat(obs))
        out['episode_window'].append([env.episode_tracker.episode_start_time_step, env.episode_tracker.episode_end_time_step])
        K = env.time_steps - 1 if steps is None else min(steps, env.time_steps - 1)
        u = rng.uniform(0.0, 1.0, size=(K, len(lo)))
        actions = (lo + u * (hi - lo)).astype('float32')
        all_actions.append(actions)
        rec = set(range(K)) if record is None else set(record(K))
        for k in range(K):
            a = [float(x) for x in actions[k]]
            if env.central_agent:
                act = [a]
            else:
                act, o = [], 0
                for s in sizes:
                    act.append(a[o:o + s])
                    o += s
            obs, rew, term, trunc, _ = env.step(act)
            if k in rec:
                out['steps'].append(k)
                out['episode'].append(ep)
                out['obs'].append(flat(obs))
                out['reward'].append(flat(rew))
                out['district'].append([float(env.net_electricity_consumption[k]), float(env.net_electricity_consumption_cost[k]),
                                        float(env.net_electricity_consumption_emission[k])])
                out['trace'].append([unit_trace(b, k) for b in env.buildings])
                out['terminated'].append(bool(term))
    arrays = {
        'actions': np.concatenate(all_actions, axis=0) if episodes == 1 else np.stack(all_actions, axis=0),
        'steps': np.array(out['steps'], dtype='int32'), 'episode': np.array(out['episode'], dtype='int32'),
        'obs': np.array(out['obs'], dtype='float32'), 'reward': np.array(out['reward'], dtype='float32'),
        'district': np.array(out['district'], dtype='float32'), 'trace': np.array(out['trace'], dtype='float32'),
        'reset_obs': np.array(out['reset_obs'], dtype='float32'), 'episode_window': np.array(out['episode_window'], dtype='int32'),
        'terminated': np.array(out['terminated'], dtype='bool'),
    }
    if env.episode_rewards:
        er = env.episode_rewards[-1]
        arrays['episode_reward_sum'] = np.array(er['sum'], dtype='float64')
        arrays['episode_reward_min'] = np.array(er['min'], dtype='float64')
        arrays['episode_reward_max'] = np.array(er['max'], dtype='float64')
    try:      # KPI table at the point where the run stopped (citylearn.py:1136-1323), for tests/test_evaluate.py
        ev = env.evaluate()
        records = [{'cost_function': r['cost_function'], 'name': r['name'], 'level': r['level'],
                    'value': None if r['value'] is None or (isinstance(r['value'], float) and np.isnan(r['value'])) else float(r['value'])}
                   for r in ev.to_dict('records')]
    except Exception as e:   # pragma: no cover
        records = [{'error': repr(e)}]
    arrays['evaluate'] = np.frombuffer(json.dumps(records).encode(), dtype='uint8')
    config = {'dataset': dataset, 'overrides': overrides or {}, 'reward': reward, 'seed': seed, 'episodes': episodes,
              'trace_names': TRACE_NAMES, 'numpy': np.__version__}
    arrays['config'] = np.frombuffer(json.dumps(config).encode(), dtype='uint8')
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype='uint8')
    OUT.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / f'{name}.npz', **arrays)
    print(name, {k: v.shape for k, v in arrays.items() if k not in ('config', 'meta')}, (OUT / f'{name}.npz').stat().st_size)


def run_ev_case(CityLearnEnv, name, dataset, overrides=None, reward=None, steps=200, seed=0, np_seed=5, hold=None, subdir='ev', episodes=1, obs_every=1):
    """Electric vehicles / chargers / washing machines.  The reference draws the SOC drift of away vehicles from NumPy's
    GLOBAL generator (citylearn/citylearn.py:1473) and a missing vehicle `initial_soc` from Python's global `random`
    (citylearn.py:2564): the fixture seeds the former (`np_seed`, = `ev_random_seed` of the replacement) and writes the latter into the
    schema (the replacement's own stable default, `citylearn_b200.ev._stable_unit`)."""
    sys.path.insert(0, str(HERE.parent))
    from citylearn_b200.ev import _stable_unit

    def hook(schema):
        for n, e in (schema.get('electric_vehicles_def') or {}).items():
            a = e['battery']['attributes']
            if a.get('initial_soc') is None:
                a['initial_soc'] = _stable_unit(n)
    np.random.seed(np_seed)
    env = make_env(CityLearnEnv, dataset, overrides, reward, schema_hook=hook)
    meta = env_meta(env)
    lo = np.concatenate([np.asarray(b.action_space.low, dtype='float64') for b in env.buildings])
    hi = np.concatenate([np.asarray(b.action_space.high, dtype='float64') for b in env.buildings])
    sizes = [b.action_space.shape[0] for b in env.buildings]
    rng = np.random.RandomState(seed)
    obs, _ = env.reset()
    for _ in range(episodes - 1):          # later episode of a rolling split: finish the earlier ones with idle actions
        while not env.terminated:
            env.step([[0.0] * n for n in sizes] if not env.central_agent else [[0.0] * sum(sizes)])
        obs, _ = env.reset()
    reset_obs = flat(obs)
    K = min(steps, env.time_steps - 1)
    actions = (lo + rng.uniform(0.0, 1.0, size=(K, len(lo))) * (hi - lo)).astype('float32')
    actions[rng.rand(*actions.shape) < 0.08] = 0.0          # exact zeros: an idle charger leaves soc[t] untouched (electric_vehicle_charger.py:325-327)
    if hold:                                                # {action name: first step with a non-zero action} - e.g. start a washing cycle late
        names = [n for b in env.buildings for n in b.active_actions]
        for an, first in hold.items():
            for j, n in enumerate(names):
                if n == an:
                    actions[:first, j] = 0.0
                    actions[first, j] = np.float32(0.5 * (lo[j] + hi[j])) if lo[j] + hi[j] != 0 else np.float32(hi[j])
    chargers = [(b, c) for b in env.buildings for c in (b.electric_vehicle_chargers or [])]
    machines = [(b, w) for b in env.buildings for w in (b.washing_machines or [])]
    out = {k: [] for k in ('obs', 'reward', 'district', 'trace', 'ev_soc', 'charger_ec', 'charger_kwh', 'wm_ec')}
    for k in range(K):
        a = [float(x) for x in actions[k]]
        act, o = [], 0
        for n in sizes:
            act.append(a[o:o + n])
            o += n
        obs, rew, term, trunc, _ = env.step([a] if env.central_agent else act)
        if k % obs_every == 0:
            out['obs'].append(flat(obs))
        out['reward'].append(flat(rew))
        out['district'].append([float(env.net_electricity_consumption[k]), float(env.net_electricity_consumption_cost[k]),
                                float(env.net_electricity_consumption_emission[k])])
        if obs_every == 1:
            out['trace'].append([unit_trace(b, k) for b in env.buildings])
        out['ev_soc'].append([float(e.battery.soc[k]) for e in env.electric_vehicles])
        out['charger_ec'].append([float(c.electricity_consumption[k]) for _, c in chargers])
        out['charger_kwh'].append([float(c.past_charging_action_values_kwh[k]) for _, c in chargers])
        out['wm_ec'].append([float(w.electricity_consumption[k]) for _, w in machines])
    arrays = {'actions': actions, 'reset_obs': np.array(reset_obs, dtype='float32'), 'obs': np.array(out['obs'], dtype='float32'),
              'reward': np.array(out['reward'], dtype='float32'), 'district': np.array(out['district'], dtype='float32'),
              'trace': np.array(out['trace'], dtype='float32'), 'ev_soc': np.array(out['ev_soc'], dtype='float32'),
              'charger_ec': np.array(out['charger_ec'], dtype='float32'), 'charger_kwh': np.array(out['charger_kwh'], dtype='float32'),
              'wm_ec': np.array(out['wm_ec'], dtype='float32')}
    config = {'dataset': dataset, 'overrides': overrides or {}, 'reward': reward, 'seed': seed, 'np_seed': np_seed, 'trace_names': TRACE_NAMES, 'episodes': episodes,
              'obs_every': obs_every,
              'episode_window': [int(env.episode_tracker.episode_start_time_step), int(env.episode_tracker.episode_end_time_step)],
              'numpy': np.__version__, 'chargers': [c.charger_id for _, c in chargers], 'vehicles': [e.name for e in env.electric_vehicles]}
    arrays['config'] = np.frombuffer(json.dumps(config).encode(), dtype='uint8')
    arrays['meta'] = np.frombuffer(json.dumps(meta).encode(), dtype='uint8')
    (OUT / subdir).mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / subdir / f'{name}.npz', **arrays)
    print(name, {k: v.shape for k, v in arrays.items() if k not in ('config', 'meta')}, (OUT / subdir / f'{name}.npz').stat().st_size)


EV_CASES = {
    'c10_evs': dict(dataset='citylearn_challenge_2022_phase_all_plus_evs', steps=300, seed=21, np_seed=5,
                    reward={'type': 'citylearn.reward_function.RewardFunction', 'attributes': {}}),
    # a washing cycle started on the LAST step of the episode (window 49..52, three profile entries, T = 52): only the entries whose
    # step lies inside the episode are added (energy_model.py:1325-1327)
    'c10_evs_short': dict(dataset='citylearn_challenge_2022_phase_all_plus_evs', steps=51, seed=23, np_seed=7,
                          overrides={'episode_time_steps': 52}, hold={'washing_machine_1': 50}),
    # charging constraints (building.py:764-989): Building_15's two chargers under a 12 kW building cap and 7 / 5 kW phase caps; actions of
    # the constrained chargers are biased towards charging so that the caps bind (scaled actions, headroom / violation observations, reward
    # penalty)
    'c11_constraints': dict(dataset='citylearn_charging_constraints_demo', steps=240, seed=31, np_seed=8),
    # CPU-side fixtures (tests/golden/ev_cpu: loader + oracle; recorded after the round's GPU budget was spent, so the GPU tests do not pick
    # them up): an episode window in the middle of the year (vehicles already plugged in / away at the first step), a central agent
    'c10_evs_window': dict(dataset='citylearn_challenge_2022_phase_all_plus_evs', steps=120, seed=24, np_seed=11, subdir='ev_cpu',
                           overrides={'simulation_start_time_step': 1200, 'simulation_end_time_step': 1500, 'episode_time_steps': 121}),
    # the whole year (8 759 steps): every connection / departure of the schedule, a year of battery degradation; observations every 24th step
    'c10_evs_year': dict(dataset='citylearn_challenge_2022_phase_all_plus_evs', steps=8759, seed=26, np_seed=13, subdir='ev_cpu', obs_every=24),
    'c10_evs_central': dict(dataset='citylearn_challenge_2022_phase_all_plus_evs', steps=100, seed=25, np_seed=12, subdir='ev_cpu',
                            overrides={'central_agent': True}),
    'c10_evs_reward': dict(dataset='citylearn_challenge_2022_phase_all_plus_evs', steps=120, seed=22, np_seed=6),
}


def run_wrapper_case(CityLearnEnv, name, dataset, wrapper, overrides=None, steps=60, seed=0):
    """Reference env under one of its own wrappers (citylearn/wrappers.py): records what the WRAPPED env returns."""
    import citylearn.wrappers as W
    base = make_env(CityLearnEnv, dataset, overrides, None)
    env = getattr(W, wrapper)(base)
    u = env.unwrapped
    space = env.action_space
    lo = np.concatenate([np.asarray(s.low, dtype='float64') for s in space])
    hi = np.concatenate([np.asarray(s.high, dtype='float64') for s in space])
    sizes = [s.shape[0] for s in space]
    rng = np.random.RandomState(seed)
    obs, _ = env.reset()
    reset_obs = flat(obs)
    K = min(steps, u.time_steps - 1)
    actions = (lo + rng.uniform(0.0, 1.0, size=(K, len(lo))) * (hi - lo)).astype('float32')
    out_obs, out_rew, out_soc = [], [], []
    for k in range(K):
        a = [float(x) for x in actions[k]]
        act, o = [], 0
        for n in sizes:
            act.append(a[o:o + n])
            o += n
        obs, rew, term, trunc, _ = env.step(act)
        out_obs.append(flat(obs))
        out_rew.append(flat(rew))
        out_soc.append([float(b.electrical_storage.soc[k]) for b in u.buildings])
    arrays = {
        'actions': actions, 'obs': np.array(out_obs, dtype='float64'), 'reward': np.array(out_rew, dtype='float32'),
        'reset_obs': np.array(reset_obs, dtype='float64'), 'soc': np.array(out_soc, dtype='float32'),
        'space_low': np.concatenate([np.asarray(s.low, dtype='float32') for s in env.observation_space]),
        'space_high': np.concatenate([np.asarray(s.high, dtype='float32') for s in env.observation_space]),
    }
    names = getattr(env, 'observation_names', u.observation_names)
    config = {'dataset': dataset, 'overrides': overrides or {}, 'wrapper': wrapper, 'seed': seed, 'observation_names': names,
              'numpy': np.__version__}
    arrays['config'] = np.frombuffer(json.dumps(config).encode(), dtype='uint8')
    (OUT / 'wrappers').mkdir(parents=True, exist_ok=True)
    np.savez_compressed(OUT / 'wrappers' / f'{name}.npz', **arrays)
    print(name, {k: v.shape for k, v in arrays.items() if k != 'config'})


WRAPPER_CASES = {
    'w_normalized_space': dict(dataset='citylearn_challenge_2022_phase_1', wrapper='NormalizedSpaceWrapper', steps=60, seed=11),
    'w_normalized_space_central': dict(dataset='citylearn_challenge_2022_phase_1', wrapper='NormalizedSpaceWrapper', overrides={'central_agent': True}, steps=40, seed=12),
    'w_normalized_obs_2023': dict(dataset='citylearn_challenge_2023_phase_2_local_evaluation', wrapper='NormalizedObservationWrapper', steps=60, seed=13),
    'w_clipped': dict(dataset='citylearn_challenge_2022_phase_1', wrapper='ClippedObservationWrapper', steps=40, seed=14),
}


CC_CONFIGS = {
    'shipped': None,                                  # the dataset's own Building_15 block
    'no_building_cap': {'Building_15': {'observations': {'headroom': True, 'violation': True, 'phase_encoding': True},
                                        'phases': [{'name': 'phase_a', 'limit_kw': 6.0, 'chargers': ['charger_15_1']},
                                                   {'name': 'phase_b', 'chargers': ['charger_15_2']}]}},
    'nothing_exposed': {'Building_15': {'building_limit_kw': 9.0, 'observations': {'headroom': False, 'violation': False, 'phase_encoding': False},
                                        'phases': [{'name': 'p1', 'limit_kw': 5.0, 'chargers': ['charger_15_1', 'charger_15_2']}]}},
    'expose_flag': {'Building_15': {'building_limit_kw': 10.0, 'expose_observations': False,
                                    'phases': [{'limit_kw': 4.0, 'chargers': ['charger_15_2']}]}},
    'unassigned': {'Building_15': {'building_limit_kw': 11.0, 'observations': {'phase_encoding': True},
                                   'phases': [{'name': 'only', 'limit_kw': 3.0, 'chargers': ['charger_15_1']}]}},
    'two_buildings': {'Building_1': {'building_limit_kw': 5.0},
                      'Building_15': {'building_limit_kw': 0.0, 'phases': [{'name': 'z', 'limit_kw': 0.0, 'chargers': ['charger_15_1']}]}},
}


def random_cc_configs(n, seed=77):
    """Random `charging_constraints` blocks for the buildings of the demo dataset that have chargers."""
    rng = np.random.RandomState(seed)
    chargers = {'Building_1': ['charger_1_1'], 'Building_4': ['charger_4_1'], 'Building_5': ['charger_5_1'], 'Building_7': ['charger_7_1'],
                'Building_10': ['charger_10_1'], 'Building_12': ['charger_12_1'], 'Building_15': ['charger_15_1', 'charger_15_2']}
    out = {}
    for i in range(n):
        cfg = {}
        names = ['Building_15'] + [b for b in chargers if b != 'Building_15' and rng.rand() < 0.35]
        for bn in names:
            ids = chargers[bn]
            c = {}
            if rng.rand() < 0.75:
                c['building_limit_kw'] = float(np.round(rng.uniform(0.0, 14.0), 2)) if rng.rand() < 0.9 else 0.0
            phases = []
            for j in range(rng.randint(0, 4)):
                ph = {'chargers': [cid for cid in ids if rng.rand() < 0.7]}
                if rng.rand() < 0.8:
                    ph['name'] = f'ph{j}' if rng.rand() < 0.8 else ''
                if rng.rand() < 0.8:
                    ph['limit_kw'] = float(np.round(rng.uniform(0.0, 9.0), 2))
                phases.append(ph)
            if phases:
                c['phases'] = phases
            obs = {}
            for key in ('headroom', 'violation', 'phase_encoding'):
                if rng.rand() < 0.6:
                    obs[key] = bool(rng.rand() < 0.6)
            if obs:
                c['observations'] = obs
            if rng.rand() < 0.2:
                c['expose_observations'] = bool(rng.rand() < 0.5)
            if not c:
                c['building_limit_kw'] = 6.0
            cfg[bn] = c
        out[f'fuzz{i}'] = cfg
    return out


def run_cc_meta(CityLearnEnv, steps=40, seed=5, np_seed=9, configs=None, out_name='cc_meta.json.gz'):
    """Charging-constraint configurations (building.py:764-833) on the demo dataset: the reference's names (both orders), spaces, and a
    short trace of observations / rewards under random actions - loader and oracle fixture `tests/golden/ev/cc_meta.json.gz`."""
    import gzip
    sys.path.insert(0, str(HERE.parent))
    from citylearn_b200.ev import _stable_unit
    cases = []
    for tag, cfg in (configs or CC_CONFIGS).items():
        def hook(schema, cfg=cfg):
            for n, e in (schema.get('electric_vehicles_def') or {}).items():
                a = e['battery']['attributes']
                if a.get('initial_soc') is None:
                    a['initial_soc'] = _stable_unit(n)
            if cfg is not None:
                for b in schema['buildings'].values():
                    b.pop('charging_constraints', None)
                for bn, c in cfg.items():
                    schema['buildings'][bn]['charging_constraints'] = c
        np.random.seed(np_seed)
        env = make_env(CityLearnEnv, 'citylearn_charging_constraints_demo', None, None, schema_hook=hook)
        lo = np.concatenate([np.asarray(b.action_space.low, dtype='float64') for b in env.buildings])
        hi = np.concatenate([np.asarray(b.action_space.high, dtype='float64') for b in env.buildings])
        sizes = [b.action_space.shape[0] for b in env.buildings]
        rng = np.random.RandomState(seed)
        obs, _ = env.reset()
        rec = {'tag': tag, 'constraints': cfg, 'np_seed': np_seed, 'observation_names': env.observation_names,
               'active_observations': [b.active_observations for b in env.buildings],
               'observation_low': [s.low.tolist() for s in env.observation_space], 'observation_high': [s.high.tolist() for s in env.observation_space],
               'reset_obs': flat(obs).astype('float32').tolist(), 'actions': [], 'obs': [], 'reward': []}
        for k in range(steps):
            a = (lo + rng.uniform(0.0, 1.0, size=len(lo)) * (hi - lo)).astype('float32')
            act, o = [], 0
            for n in sizes:
                act.append([float(x) for x in a[o:o + n]])
                o += n
            obs, rew, _, _, _ = env.step(act)
            rec['actions'].append(a.tolist()); rec['obs'].append(flat(obs).astype('float32').tolist()); rec['reward'].append(flat(rew).astype('float32').tolist())
        cases.append(rec)
        print('cc_meta', tag, len(rec['observation_names'][14]))
    (OUT / 'ev').mkdir(parents=True, exist_ok=True)
    with gzip.open(OUT / 'ev' / out_name, 'wt') as f:
        json.dump(cases, f)
    print(out_name, len(cases), (OUT / 'ev' / out_name).stat().st_size)


def run_meta_fuzz(CityLearnEnv, n=28, seed=123):
    """Loader fuzz: random combinations of the constructor overrides -> the reference's names, spaces and episode windows."""
    rng = np.random.RandomState(seed)
    datasets = [P1, PALL, C23, Z20, 'citylearn_challenge_2023_phase_3_1', 'citylearn_challenge_2022_phase_3']
    cases = []
    for i in range(n):
        ds = datasets[i % len(datasets)]
        sch = json.load(open(DATASETS / ds / 'schema.json'))
        names = list(sch['buildings'])
        obs_names = [k for k, v in sch['observations'].items() if v['active']]
        ov = {}
        if rng.rand() < 0.5:
            ov['central_agent'] = bool(rng.rand() < 0.5)
        if rng.rand() < 0.5:
            k = rng.randint(1, len(names) + 1)
            ov['buildings'] = sorted(rng.choice(len(names), size=k, replace=False).tolist()) if rng.rand() < 0.5 else \
                [names[j] for j in sorted(rng.choice(len(names), size=k, replace=False).tolist())]
        if rng.rand() < 0.5:
            ov['inactive_observations'] = [obs_names[j] for j in sorted(rng.choice(len(obs_names), size=rng.randint(1, 5), replace=False).tolist())]
        end = sch['simulation_end_time_step']
        if rng.rand() < 0.6:
            a = int(rng.randint(0, end // 2)); b = int(rng.randint(a + 48, end + 1))
            ov['simulation_start_time_step'], ov['simulation_end_time_step'] = a, b
            if rng.rand() < 0.7:
                ov['episode_time_steps'] = int(rng.randint(24, max(25, (b - a) // 2)))
                ov['rolling_episode_split'] = bool(rng.rand() < 0.5)
        if rng.rand() < 0.3:
            ov['shared_observations'] = ['month', 'hour'] if 'month' in obs_names else ['hour']
        try:
            env = make_env(CityLearnEnv, ds, ov, None)
            windows = []
            for ep in range(3):
                env.reset()
                windows.append([env.episode_tracker.episode_start_time_step, env.episode_tracker.episode_end_time_step])
        except Exception as e:      # e.g. ReliabilityMetricsPowerOutage.get_signals breaks on some episode lengths (power_outage.py:154)
            print('meta', i, ds, ov, 'reference raised', repr(e)[:80])
            continue
        cases.append({'dataset': ds, 'overrides': ov, 'observation_names': env.observation_names, 'action_names': env.action_names,
                      'observation_low': [s.low.tolist() for s in env.observation_space], 'observation_high': [s.high.tolist() for s in env.observation_space],
                      'action_low': [s.low.tolist() for s in env.action_space], 'action_high': [s.high.tolist() for s in env.action_space],
                      'central_agent': env.central_agent, 'shared_observations': env.shared_observations, 'windows': windows,
                      'building_names': [b.name for b in env.buildings]})
        print('meta', i, ds, ov)
    with open(OUT / 'meta_fuzz.json', 'w') as f:
        json.dump({'numpy': np.__version__, 'cases': cases}, f)


def run_trace_fuzz(CityLearnEnv, seed=321, plan=None, out_name='trace_fuzz.json.gz', sub_windows=True):
    """Physics fuzz: short reference runs under random override combinations (sub-windows, building subsets, central agent, rewards);
    only observations / rewards / district sums are stored."""
    rng = np.random.RandomState(seed)
    rewards = [None, MARL, {'type': 'citylearn.reward_function.IndependentSACReward', 'attributes': {}},
               {'type': 'citylearn.reward_function.SolarPenaltyReward', 'attributes': {}},
               {'type': 'citylearn.reward_function.RewardFunction', 'attributes': {'exponent': 1.5}}]
    plan = plan or [(P1, 40), (PALL, 30), (C23, 80), (Z20, 40), ('citylearn_challenge_2023_phase_3_1', 80), ('citylearn_challenge_2022_phase_3', 40),
                    (C23, 60), (Z20, 40), (PALL, 30), ('citylearn_challenge_2023_phase_1', 80), ('citylearn_challenge_2020_climate_zone_3', 40), (P1, 40)]
    cases = []
    for i, (ds, steps) in enumerate(plan):
        sch = json.load(open(DATASETS / ds / 'schema.json'))
        names = list(sch['buildings'])
        ov = {'central_agent': bool(rng.rand() < 0.5)}
        if rng.rand() < 0.6 and len(names) > 2:
            k = rng.randint(2, len(names) + 1)
            ov['buildings'] = [names[j] for j in sorted(rng.choice(len(names), size=k, replace=False).tolist())]
        end = sch['simulation_end_time_step']
        if sub_windows:
            a = int(rng.randint(0, end // 2))
            ov['simulation_start_time_step'], ov['simulation_end_time_step'] = a, int(min(end, a + rng.randint(steps + 30, steps + 400)))
        reward = rewards[rng.randint(len(rewards))]
        if 'LSTM' in json.dumps(sch['buildings'][names[0]].get('type', '')) and reward is not None and 'SolarPenalty' in reward['type']:
            reward = None
        try:
            env = make_env(CityLearnEnv, ds, ov, reward)
            obs, _ = env.reset()
        except Exception as e:
            print('trace', i, ds, ov, 'reference raised', repr(e)[:80])
            continue
        lo = np.concatenate([np.asarray(b.action_space.low, dtype='float64') for b in env.buildings])
        hi = np.concatenate([np.asarray(b.action_space.high, dtype='float64') for b in env.buildings])
        sizes = [b.action_space.shape[0] for b in env.buildings]
        K = min(steps, env.time_steps - 1)
        actions = (lo + rng.uniform(size=(K, len(lo))) * (hi - lo)).astype('float32')
        rec = {'dataset': ds, 'overrides': ov, 'reward': reward, 'reset_obs': [float(np.float32(v)) for v in flat(obs)], 'actions': actions.tolist(), 'obs': [], 'reward_values': [], 'district': []}
        failed = None
        for k in range(K):
            a_ = [float(x) for x in actions[k]]
            if env.central_agent:
                act = [a_]
            else:
                act, o = [], 0
                for n_ in sizes:
                    act.append(a_[o:o + n_]); o += n_
            try:
                obs, rew, term, _, _ = env.step(act)
            except AssertionError as e:       # the reference's own limit checks (e.g. t = 0 double counting vs an autosized device)
                failed = repr(e)[:80]
                break
            rec['obs'].append([float(np.float32(v)) for v in flat(obs)])
            rec['reward_values'].append([float(np.float32(v)) for v in flat(rew)])
            rec['district'].append([float(np.float32(env.net_electricity_consumption[k])), float(np.float32(env.net_electricity_consumption_cost[k])),
                                    float(np.float32(env.net_electricity_consumption_emission[k]))])
        if failed:
            print('trace', i, ds, ov, 'reference asserted', failed)
            continue
        cases.append(rec)
        print('trace', i, ds, ov, None if reward is None else reward['type'].split('.')[-1], K)
    import gzip
    with gzip.open(OUT / out_name, 'wt') as f:
        json.dump({'numpy': np.__version__, 'cases': cases}, f)


def sparse(K):
    return sorted(set(list(range(0, 48)) + list(range(0, K, 41)) + list(range(K - 48, K))))


P1 = 'citylearn_challenge_2022_phase_1'
PALL = 'citylearn_challenge_2022_phase_all'
C23 = 'citylearn_challenge_2023_phase_2_local_evaluation'
Z20 = 'citylearn_challenge_2020_climate_zone_1'
BAEDA = 'baeda_3dem'
C23P3 = 'citylearn_challenge_2023_phase_3_1'
MARL = {'type': 'citylearn.reward_function.MARL', 'attributes': {}}

CASES = {
    # C1: 5 buildings, decentralised, default reward (BASELINE.json configs[0])
    'c1_phase1_300': dict(dataset=P1, steps=300),
    'c1_phase1_central': dict(dataset=P1, overrides={'central_agent': True}, steps=60),
    # C2 schema: 17 buildings, full year, sparse recording + episode reward sums (SURVEY.md Appendix D.2 uses its own actions)
    'c2_year': dict(dataset=PALL, steps=None, record=sparse),
    'c2_marl': dict(dataset=PALL, reward=MARL, steps=120),
    'c2_isac': dict(dataset=PALL, reward={'type': 'citylearn.reward_function.IndependentSACReward', 'attributes': {}}, steps=60),
    'c2_solar_penalty': dict(dataset=PALL, reward={'type': 'citylearn.reward_function.SolarPenaltyReward', 'attributes': {}}, steps=120),
    'c2_central_exp2': dict(dataset=PALL, overrides={'central_agent': True},
                            reward={'type': 'citylearn.reward_function.RewardFunction', 'attributes': {'exponent': 2.0}}, steps=60),
    # per-building reward functions (MultiBuildingRewardFunction, citylearn.py:2106-2141): named entries + the 'default' fallback
    'c1_multi_reward': dict(dataset=P1, reward={'type': {'Building_1': 'citylearn.reward_function.RewardFunction',
                                                         'Building_2': 'citylearn.reward_function.IndependentSACReward',
                                                         'default': 'citylearn.reward_function.SolarPenaltyReward'},
                                                'attributes': {'Building_1': {'exponent': 2.0}, 'default': {}}}, steps=80, seed=23),
    # episode windows: consecutive splits inside a sub-range, two episodes
    'c1_episodes': dict(dataset=P1, overrides={'simulation_start_time_step': 100, 'simulation_end_time_step': 1299,
                                               'episode_time_steps': 240}, episodes=3, seed=3),
    # sub-hour control step on an hourly dataset: time_step_ratio = 0.5 (tests/unit/test_subhour_scaling.py in the reference)
    'c1_subhour': dict(dataset=P1, overrides={'seconds_per_time_step': 1800}, steps=120, seed=5),
    # C3 schema: 3 LSTM buildings, outages, decentralised MARL (BASELINE.json configs[2]) and the schema default
    'c3_marl': dict(dataset=C23, overrides={'central_agent': False}, reward=MARL, steps=None, seed=1),
    'c3_default_central_comfort': dict(dataset=C23, steps=None, seed=2),
    'c3_solar_comfort': dict(dataset=C23, overrides={'central_agent': False},
                             reward={'type': 'citylearn.reward_function.SolarPenaltyAndComfortReward',
                                     'attributes': {'band': 1.0, 'lower_exponent': 2.0, 'higher_exponent': 3.0, 'coefficients': [1.0, 2.0]}},
                             steps=200, seed=4),
    # thermal tanks + autosized heat pumps / heaters / tanks (2020 challenge, 9 buildings; cooling, dhw and battery actions)
    'c6_tanks_2020': dict(dataset=Z20, steps=None, record=sparse, seed=6),
    'c6_tanks_2020_marl_central': dict(dataset=Z20, overrides={'central_agent': True}, reward=MARL, steps=150, seed=7),
    'c6_tanks_2020_solar_penalty': dict(dataset=Z20, reward={'type': 'citylearn.reward_function.SolarPenaltyReward', 'attributes': {}}, steps=150, seed=16),
    # cooling tank + cooling-device action on LSTM buildings (hidden 8, 11 inputs); Building_4's 1x50 LSTM is outside the kernel's shape
    'c6_baeda3': dict(dataset=BAEDA, overrides={'buildings': ['Building_1', 'Building_2', 'Building_3']}, steps=600, seed=8),
    # 32-building slice of the synthetic wide district (C4): pins the per-building path of the 1024-building runs
    'c4_slice32': dict(dataset='synthetic_wide_32', steps=200, seed=10),
    # 2020 district with a synthetic heating season: heating heat pump + heating tank + the tank-capacity quirks (no bundled dataset has them)
    'c8_heating': dict(dataset='synthetic_heating', steps=400, seed=18),
    'c8_heating_central_marl': dict(dataset='synthetic_heating', overrides={'central_agent': True}, reward=MARL, steps=120, seed=19),
    # LSTM buildings driven through the signed cooling_or_heating_device action, heating + auto hvac modes
    'c9_dual_mode': dict(dataset='synthetic_dual_mode', steps=500, seed=20),
    # 6 LSTM buildings with stochastic outages, central agent, full 2207-step episode
    'c7_phase3': dict(dataset=C23P3, steps=None, record=sparse, seed=9),
}

if __name__ == '__main__':
    CityLearnEnv = import_reference()
    todo = sys.argv[1:] or (list(CASES) + list(EV_CASES) + list(WRAPPER_CASES) + ['meta_fuzz', 'trace_fuzz', 'trace_datasets', 'cc_meta', 'cc_fuzz'])
    for n in todo:
        if n == 'trace_fuzz':
            run_trace_fuzz(CityLearnEnv)
            continue
        if n == 'trace_datasets':        # every other bundled dataset family from its first time step (full windows)
            run_trace_fuzz(CityLearnEnv, seed=77, out_name='trace_datasets.json.gz', sub_windows=False, plan=[
                ('citylearn_challenge_2021', 60), ('citylearn_challenge_2020_climate_zone_2', 60), ('citylearn_challenge_2020_climate_zone_4', 60),
                ('citylearn_challenge_2022_phase_2', 60), ('citylearn_challenge_2023_phase_2_online_evaluation_1', 60),
                ('citylearn_challenge_2023_phase_2_online_evaluation_2', 60), ('citylearn_challenge_2023_phase_2_online_evaluation_3', 60),
                ('citylearn_challenge_2023_phase_3_2', 60), ('citylearn_challenge_2023_phase_3_3', 60)])
            continue
        if n == 'cc_meta':
            run_cc_meta(CityLearnEnv)
            continue
        if n == 'cc_fuzz':
            run_cc_meta(CityLearnEnv, steps=30, seed=6, np_seed=10, configs=random_cc_configs(12), out_name='cc_fuzz.json.gz')
            continue
        if n == 'meta_fuzz':
            run_meta_fuzz(CityLearnEnv)
            continue
        if n in EV_CASES:
            run_ev_case(CityLearnEnv, n, **EV_CASES[n])
        elif n in WRAPPER_CASES:
            run_wrapper_case(CityLearnEnv, n, **WRAPPER_CASES[n])
        else:
            run_case(CityLearnEnv, n, **CASES[n])
