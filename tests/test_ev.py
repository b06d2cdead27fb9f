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


def oracle_reset(env, spec):
    tracker = S.EpisodeTracker(spec.simulation_start_time_step, spec.simulation_end_time_step)
    ets = spec.episode_time_steps if spec.episode_time_steps is not None else tracker.simulation_time_steps
    tracker.next_episode(ets, spec.rolling_episode_split, spec.random_episode_split, spec.random_seed)
    return env.reset(tracker.episode_start_time_step, tracker.episode_time_steps)


@pytest.mark.parametrize('case', CASES)
def test_loader_reproduces_reference_names_and_spaces(case):
    z, cfg, meta, spec = load(case)
    # env.observation_names are the keys of Building.observations() (value order); the space follows active_observations - the two orders
    # differ for a building with charging constraints (reference quirk, kept)
    assert [list(b.observation_value_order or b.active_observations) for b in spec.buildings] == meta['observation_names']
    assert [list(b.active_observations) for b in spec.buildings] == [m['active_observations'] for m in meta['buildings']]
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
    assert np.array_equal(oracle_reset(env, spec)[0].astype('float32'), z['reset_obs'])
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


# ---------------------------------------------------------------------------------------------------------------- GPU (C ABI) parity
def make_gpu_env(cfg, **kw):
    from citylearn_b200 import CityLearnEnv
    src = DataSet.get_source(cfg['dataset'])
    sch = src.schema()
    if cfg.get('reward') is not None:
        sch['reward_function'] = {'type': cfg['reward']['type'], 'attributes': cfg['reward'].get('attributes', {})}
    return CityLearnEnv(sch, data_source=src, ev_random_seed=cfg['np_seed'], **(cfg.get('overrides') or {}), **kw)


def ev_soc_entries(env):
    import torch
    n = env.spec.ev['n_ev']
    prev = torch.zeros((env.num_envs, n), dtype=torch.float32, device=env.device)
    now = torch.zeros_like(prev)
    env._h.ev_read(prev.data_ptr(), now.data_ptr(), torch.cuda.current_stream(env.device).cuda_stream)
    torch.cuda.synchronize()
    return prev.cpu().numpy(), now.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_gpu_single_env_matches_reference_bit_for_bit(case):
    """The step kernel (fp64 flow) against the recorded reference run: observations, rewards, district sums and every vehicle's SOC
    entry at every step, through `CityLearnEnv.step` with nested-list actions (the reference's call)."""
    z, cfg, meta, spec = load(case)
    env = make_gpu_env(cfg, num_envs=1)
    obs, _ = env.reset()
    assert np.array_equal(np.array([v for row in obs for v in row], dtype='float32'), z['reset_obs'])
    sizes = [len(b.active_actions) for b in env.spec.buildings]
    central = env.central_agent
    for k in range(len(z['actions'])):
        a = [float(x) for x in z['actions'][k]]
        nested, o = [], 0
        for s in sizes:
            nested.append(a[o:o + s])
            o += s
        obs, rew, term, trunc, _ = env.step([a] if central else nested)
        assert np.array_equal(np.array([v for row in obs for v in row], dtype='float32'), z['obs'][k]), f'obs step {k}'
        r = np.array(rew, dtype='float32')
        # a central agent's reward is the float32 sum of the buildings' rewards in building order, like the reference's
        assert np.array_equal(r, z['reward'][k]), f'reward step {k}: {np.abs(r - z["reward"][k]).max()}'
        assert np.array_equal(env.district[0].cpu().numpy(), z['district'][k]), f'district step {k}'
        prev, now = ev_soc_entries(env)
        assert np.array_equal(prev[0], z['ev_soc'][k]), f'vehicle soc step {k}'
    assert env.gpu_launches > 0


@pytest.mark.gpu
@pytest.mark.parametrize('case,precision', [('c10_evs_reward', 'fp64'), ('c10_evs_reward', 'fp32'), ('c11_constraints', 'fp64')])
def test_gpu_batched_envs_match_oracle(case, precision):
    """96 envs with different action sequences (exact zeros included), 160 steps (Electric_Vehicles_Reward_Function; with and without
    charging constraints - scaled actions, headroom / violation observations, reward penalty), against the vectorised oracle:
    bit-exact in the fp64 flow; fp32: 1e-4 scaled for observations / district sums / SOCs, threshold flips of the reward's step terms
    in under 0.1 % of the entries; one `rollout` launch equals the step-by-step run."""
    import torch
    z, cfg, meta, spec = load(case)
    E, K = 96, 160
    env = make_gpu_env(cfg, num_envs=E, precision=precision)
    ora = OracleEnv(spec, E, libm_pow=False)
    rng = np.random.RandomState(3)
    lo = np.concatenate([b.action_low for b in spec.buildings]).astype('float64')
    hi = np.concatenate([b.action_high for b in spec.buildings]).astype('float64')
    acts = (lo + rng.uniform(size=(K, E, lo.size)) * (hi - lo)).astype('float32')
    acts[rng.rand(*acts.shape) < 0.1] = 0.0
    o0 = oracle_reset(ora, spec)
    assert np.array_equal(env.observations.cpu().numpy(), o0.astype('float32'))
    roll = make_gpu_env(cfg, num_envs=E, precision=precision)
    ro = torch.zeros((K, E, roll._obs_dim), device='cuda')
    rr = torch.zeros((K, E, roll._reward_dim), device='cuda')
    rd = torch.zeros((K, E, 3), device='cuda')
    roll.rollout(torch.as_tensor(acts, device='cuda'), ro, rr, rd)
    worst, flips, total = 0.0, 0, 1
    for k in range(K):
        obs, rew, term, _, _ = env.step(torch.as_tensor(acts[k], device='cuda'))
        oo, orw, od, _ = ora.step(acts[k])
        got = {'obs': obs.cpu().numpy(), 'reward': rew.cpu().numpy(), 'district': env.district.cpu().numpy(), 'soc': ev_soc_entries(env)[0]}
        ref = {'obs': oo.astype('float32'), 'reward': orw.astype('float32').reshape(got['reward'].shape), 'district': od.astype('float32'),
               'soc': ora.ev_soc_prev.astype('float32')}
        for name in got:
            if precision == 'fp64':
                assert np.array_equal(got[name], ref[name]), f'{name} step {k}: {np.abs(got[name] - ref[name]).max()}'
            else:
                err = np.abs(got[name] - ref[name]) / np.maximum(np.abs(ref[name]), 1.0)
                if name == 'reward':          # the charger terms are step functions of the SOC: float32 rounding may flip one
                    flips += int((err > 1e-3).sum()); total += err.size
                else:
                    worst = max(worst, float(err.max()))
        assert torch.equal(ro[k], obs) and torch.equal(rr[k], rew) and torch.equal(rd[k], env.district), f'rollout step {k}'
    assert worst <= 1e-4 and flips <= 1e-3 * total, (worst, flips, total)


@pytest.mark.gpu
def test_gpu_checkpoint_resume_and_second_episode():
    """state_dict / load_state_dict carry the vehicles' state; a second episode starts from the same initial SOCs."""
    import torch
    z, cfg, meta, spec = load('c10_evs')
    E = 8
    a = make_gpu_env(cfg, num_envs=E, episode_time_steps=48)
    b = make_gpu_env(cfg, num_envs=E, episode_time_steps=48)
    g = torch.Generator().manual_seed(1)
    acts = (torch.rand((47, E, a.spec.action_dim), generator=g) * 2 - 1).cuda()
    first = []
    for k in range(20):
        first.append(a.step(acts[k])[1].clone())
    sd = a.state_dict()
    tail = [a.step(acts[k])[1].clone() for k in range(20, 47)]
    assert a.terminated
    b.load_state_dict(sd)
    for k in range(20, 47):
        assert torch.equal(b.step(acts[k])[1], tail[k - 20]), k
    assert torch.equal(b.observations, a.observations)
    assert np.array_equal(ev_soc_entries(a)[1], ev_soc_entries(b)[1])
    a.reset()
    b2 = make_gpu_env(cfg, num_envs=E, episode_time_steps=48)
    b2.reset()                                    # rolling episodes: same second window
    for k in range(10):
        assert torch.equal(a.step(acts[k])[1], b2.step(acts[k])[1]), k


@pytest.mark.gpu
def test_gpu_unsupported_combinations_fail_loudly():
    import torch
    z, cfg, meta, spec = load('c10_evs')
    with pytest.raises(NotImplementedError):
        make_gpu_env(cfg, num_envs=4, stale_observations=False)
    with pytest.raises(NotImplementedError):
        make_gpu_env(cfg, num_envs=4, track_kpis=True)
    with pytest.raises(NotImplementedError):
        make_gpu_env(cfg, num_envs=4, central_agent=True)
    env = make_gpu_env(cfg, num_envs=4, episode_time_steps=48)
    with pytest.raises(NotImplementedError):
        env.reset(options={'episode_start': torch.tensor([0, 24, 48, 72])})
    with pytest.raises(RuntimeError):
        env.evaluate()
    from citylearn_b200 import CityLearnEnv
    from citylearn_b200.reward_function import Electric_Vehicles_Reward_Function
    with pytest.raises(ValueError, match='chargers'):
        CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=2, reward_function=Electric_Vehicles_Reward_Function)


# ------------------------------------------------------------------------------------------ charging-constraint configurations
def _cc_cases():
    import gzip
    p = GOLDEN / 'ev' / 'cc_meta.json.gz'
    return json.load(gzip.open(p, 'rt')) if p.exists() else []


CC_CASES = _cc_cases()


def _cc_spec(rec):
    src = DataSet.get_source('citylearn_charging_constraints_demo')
    sch = src.schema()
    if rec['constraints'] is not None:
        for b in sch['buildings'].values():
            b.pop('charging_constraints', None)
        for bn, c in rec['constraints'].items():
            sch['buildings'][bn]['charging_constraints'] = c
    return S.load(sch, data_source=src, ev_random_seed=rec['np_seed'])


@pytest.mark.parametrize('rec', CC_CASES, ids=[r['tag'] for r in CC_CASES])
def test_charging_constraint_configurations_match_the_reference(rec):
    """Six `charging_constraints` blocks run through the unmodified reference (oracle/make_golden.py cc_meta): caps with and without a
    building limit, a phase without a cap, nothing exposed, `expose_observations`, an unassigned charger in the phase encoding, zero
    caps on two buildings.  The loader reproduces both name orders and the spaces; the oracle the observations and rewards of 40 steps."""
    spec = _cc_spec(rec)
    assert [list(b.observation_value_order or b.active_observations) for b in spec.buildings] == rec['observation_names']
    assert [list(b.active_observations) for b in spec.buildings] == rec['active_observations']
    for b, lo, hi in zip(spec.buildings, rec['observation_low'], rec['observation_high']):
        assert np.array_equal(b.observation_low, np.float32(lo)) and np.array_equal(b.observation_high, np.float32(hi)), b.name
    env = OracleEnv(spec, 1, libm_pow=True)
    assert np.array_equal(oracle_reset(env, spec)[0].astype('float32'), np.float32(rec['reset_obs']))
    for k, a in enumerate(rec['actions']):
        obs, rew, _, _ = env.step(np.float32(a)[None])
        assert np.array_equal(obs[0], np.float32(rec['obs'][k])), k
        assert np.array_equal(rew[0], np.float32(rec['reward'][k])), k


@pytest.mark.gpu
@pytest.mark.parametrize('rec', CC_CASES, ids=[r['tag'] for r in CC_CASES])
def test_gpu_charging_constraint_configurations_match_the_reference(rec):
    from citylearn_b200 import CityLearnEnv
    env = CityLearnEnv(_cc_spec(rec), num_envs=1)
    obs, _ = env.reset()
    assert np.array_equal(np.array([v for row in obs for v in row], dtype='float32'), np.float32(rec['reset_obs']))
    sizes = [len(b.active_actions) for b in env.spec.buildings]
    for k, a in enumerate(rec['actions']):
        nested, o = [], 0
        for s in sizes:
            nested.append([float(x) for x in a[o:o + s]])
            o += s
        obs, rew, _, _, _ = env.step(nested)
        assert np.array_equal(np.array([v for row in obs for v in row], dtype='float32'), np.float32(rec['obs'][k])), k
        assert np.array_equal(np.array(rew, dtype='float32'), np.float32(rec['reward'][k])), k


def _cc_fuzz_cases():
    import gzip
    p = GOLDEN / 'ev' / 'cc_fuzz.json.gz'
    return json.load(gzip.open(p, 'rt')) if p.exists() else []


CC_FUZZ = _cc_fuzz_cases()


@pytest.mark.parametrize('rec', CC_FUZZ, ids=[r['tag'] for r in CC_FUZZ])
def test_random_charging_constraint_blocks_match_the_reference(rec):
    """Twelve random `charging_constraints` blocks (1 - 4 constrained buildings, 0 - 3 phases with and without names / caps / members,
    random observation flags) through the unmodified reference, 30 steps each: loader names (both orders), spaces, and the oracle's
    observations and rewards."""
    try:
        spec = _cc_spec(rec)
    except S.UnsupportedSchemaError:
        pytest.skip('more phases / phase members than the device supports')
    assert [list(b.observation_value_order or b.active_observations) for b in spec.buildings] == rec['observation_names']
    assert [list(b.active_observations) for b in spec.buildings] == rec['active_observations']
    for b, lo, hi in zip(spec.buildings, rec['observation_low'], rec['observation_high']):
        assert np.array_equal(b.observation_low, np.float32(lo)) and np.array_equal(b.observation_high, np.float32(hi)), b.name
    env = OracleEnv(spec, 1, libm_pow=True)
    assert np.array_equal(oracle_reset(env, spec)[0].astype('float32'), np.float32(rec['reset_obs']))
    for k, a in enumerate(rec['actions']):
        obs, rew, _, _ = env.step(np.float32(a)[None])
        assert np.array_equal(obs[0], np.float32(rec['obs'][k])), k
        assert np.array_equal(rew[0], np.float32(rec['reward'][k])), k


CPU_CASES = sorted(p.stem for p in (GOLDEN / 'ev_cpu').glob('*.npz'))


@pytest.mark.parametrize('case', CPU_CASES)
def test_oracle_matches_reference_on_windows_and_central_agent(case):
    """An episode window in the middle of the year (vehicles plugged in / away at its first step: the episode-start SOC columns), a
    central agent (shared observations dropped, one summed reward) and the WHOLE year (8 759 steps: every connection / departure of the
    schedule, a year of vehicle-battery degradation) - loader and oracle against the unmodified reference."""
    z = np.load(GOLDEN / 'ev_cpu' / f'{case}.npz')
    cfg = json.loads(bytes(z['config']).decode())
    meta = json.loads(bytes(z['meta']).decode())
    src = DataSet.get_source(cfg['dataset'])
    sch = src.schema()
    if cfg.get('reward') is not None:
        sch['reward_function'] = {'type': cfg['reward']['type'], 'attributes': cfg['reward'].get('attributes', {})}
    spec = S.load(sch, data_source=src, ev_random_seed=cfg['np_seed'], **(cfg.get('overrides') or {}))
    env = OracleEnv(spec, 1, libm_pow=True)
    assert env.entries is not None
    names = [[n for bi2, n in env.entries if bi2 == bi] for bi in range(len(spec.buildings))] if not spec.central_agent else [[n for _, n in env.entries]]
    assert names == meta['observation_names']
    assert np.array_equal(oracle_reset(env, spec)[0].astype('float32'), z['reset_obs'])
    assert [int(env.start[0]), int(env.start[0]) + env.T - 1] == cfg['episode_window']
    every = int(cfg.get('obs_every', 1))                 # (the full-year run keeps the observations of every 24th step)
    for k in range(len(z['actions'])):
        obs, rew, dist, dyn = env.step(z['actions'][k][None])
        if k % every == 0:
            assert np.array_equal(obs[0], z['obs'][k // every]), k
        assert np.array_equal(dist[0], z['district'][k]), k
        assert np.array_equal(env.ev_soc_prev[0].astype('float32'), z['ev_soc'][k]), k
        r = np.asarray(rew[0], dtype='float32').reshape(-1)
        ok = np.array_equal(r, z['reward'][k]) if not spec.central_agent else bool(np.abs(r - z['reward'][k]).max() <= 2e-6 * max(1.0, float(np.abs(z['reward'][k]).max())))
        assert ok, (k, r, z['reward'][k])
