"""Fused wrappers (citylearn_b200/wrappers.py, cl_set_transforms) vs the reference's own wrappers (tests/golden/wrappers/*.npz,
recorded by oracle/make_golden.py from citylearn/wrappers.py)."""
import json

import numpy as np
import pytest

from citylearn_b200 import schema as S
from citylearn_oracle import OracleEnv
from helpers import GOLDEN, schema_for

CASES = sorted(p.stem for p in (GOLDEN / 'wrappers').glob('*.npz'))
MODE = {'NormalizedSpaceWrapper': ('normalized', True), 'NormalizedObservationWrapper': ('normalized', False),
        'ClippedObservationWrapper': ('clipped', False)}


def load(case):
    z = np.load(GOLDEN / 'wrappers' / f'{case}.npz')
    return z, json.loads(bytes(z['config']).decode())


def apply_transform(t, v):
    """cl_obs_transform semantics in float64 (include/citylearn_b200.h)."""
    v = np.asarray(v, dtype='float64')
    w = t['w'].astype('float64')
    v = np.where(t['fn'] == S.OBS_FN_SIN, np.sin(v * w), np.where(t['fn'] == S.OBS_FN_COS, np.cos(v * w), v))
    v = v * t['scale'].astype('float64') + t['offset'].astype('float64')
    return np.clip(v, t['lo'], t['hi'])


@pytest.mark.parametrize('case', CASES)
def test_transform_layout_reproduces_reference_wrappers(case):
    """Host side: names, expanded layout and the per-column transform records, applied to the oracle's raw observations."""
    z, cfg = load(case)
    sch, src, ov = schema_for(cfg)
    spec = S.load(sch, data_source=src, **ov)
    mode, norm_actions = MODE[cfg['wrapper']]
    entries, desc = S.observation_layout(spec)
    out_entries, out_desc, t = S.transformed_observation_layout(spec, entries, desc, mode)
    names = [n for row in cfg['observation_names'] for n in row]
    assert [n for _, n in out_entries] == names
    if mode == 'normalized':
        assert np.all(z['space_low'] == 0.0) and np.all(z['space_high'] == 1.0)
    src_of = []                      # column of the raw row every output column reads
    j = 0
    for k, (bi, n) in enumerate(out_entries):
        while entries[j][0] != bi or not (n == entries[j][1] or n in (entries[j][1] + '_cos', entries[j][1] + '_sin')):
            j += 1
        src_of.append(j)
    env = OracleEnv(spec, 1)
    raw = env.reset()[0]
    np.testing.assert_allclose(apply_transform(t, raw[src_of]), z['reset_obs'], rtol=0, atol=2e-6)
    lo = np.array([v for b in spec.buildings for v in b.action_low], dtype='float32')
    hi = np.array([v for b in spec.buildings for v in b.action_high], dtype='float32')
    for k in range(len(z['actions'])):
        a = z['actions'][k]
        a = (a * (hi - lo) + lo).astype('float32') if norm_actions else a       # wrappers.py:208-222, float32 arithmetic
        obs, rew, dist, dyn = env.step(a[None])
        np.testing.assert_allclose(apply_transform(t, obs[0][src_of]), z['obs'][k], rtol=0, atol=2e-6)
        np.testing.assert_allclose(rew[0], z['reward'][k], rtol=1e-5, atol=1e-6)
        # NormalizedActionWrapper hands np.float32 actions to the reference env, whose `action * nominal_power` then runs in float32
        # instead of float64 (NumPy weak-scalar promotion): the reference itself moves by an ulp under its own wrapper
        soc = dyn[0, :, S.DYN['electrical_storage_soc']].astype('float32')
        if norm_actions:
            np.testing.assert_allclose(soc, z['soc'][k], rtol=0, atol=3e-7)
        else:
            assert np.array_equal(soc, z['soc'][k])


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_fused_wrappers_match_reference(case):
    """Device side: the wrapped env returns the reference wrapper's values (float32 sin / cos / affine: 2e-6 absolute on [0, 1])."""
    from citylearn_b200 import CityLearnEnv, wrappers as W
    z, cfg = load(case)
    sch, src, ov = schema_for(cfg)
    base = CityLearnEnv(sch, data_source=src, num_envs=1, **ov)
    env = getattr(W, cfg['wrapper'])(base)
    assert [n for row in (env.observation_names) for n in row] == [n for row in cfg['observation_names'] for n in row]
    assert np.array_equal(np.concatenate([s.low for s in env.observation_space]), z['space_low'])
    assert np.array_equal(np.concatenate([s.high for s in env.observation_space]), z['space_high'])
    obs, _ = env.reset()
    np.testing.assert_allclose(np.array([v for row in obs for v in row]), z['reset_obs'], rtol=0, atol=2e-6)
    sizes = [len(r) for r in env.unwrapped.action_names]
    for k in range(len(z['actions'])):
        a = [float(x) for x in z['actions'][k]]
        nested, o = [], 0
        for n in sizes:
            nested.append(a[o:o + n])
            o += n
        obs, rew, term, _, _ = env.step(nested)
        np.testing.assert_allclose(np.array([v for row in obs for v in row]), z['obs'][k], rtol=0, atol=2e-6)
        np.testing.assert_allclose(np.array(rew, dtype='float32'), z['reward'][k], rtol=1e-5, atol=1e-5)   # LSTM comfort rewards: 1e-5 of scale 1
    # the unwrapped env still reports the raw spaces
    assert len(np.concatenate([s.low for s in env.unwrapped.observation_space])) == sum(len(b.active_observations) for b in base.spec.buildings) \
        or base.central_agent


@pytest.mark.gpu
def test_fused_wrappers_batched_rollout_and_sb3():
    """Batched envs: normalised observations are the affine image of the raw ones; SB3 wrapper shapes on tensors."""
    import torch
    from citylearn_b200 import CityLearnEnv, wrappers as W
    E, K = 32, 20
    raw = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E, central_agent=True)
    wrapped = W.StableBaselines3Wrapper(W.NormalizedSpaceWrapper(CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E, central_agent=True)))
    assert wrapped.observation_space.shape == (wrapped.unwrapped._obs_dim,)
    lo = torch.tensor([v for b in raw.spec.buildings for v in b.action_low], device='cuda')
    hi = torch.tensor([v for b in raw.spec.buildings for v in b.action_high], device='cuda')
    g = torch.Generator(device='cuda').manual_seed(3)
    raw.reset(); wrapped.reset()
    u = wrapped.unwrapped
    t = S.transformed_observation_layout(u.spec, u._entries, u._raw_desc, 'normalized')[2]
    names_raw = [n for _, n in raw._entries]
    for k in range(K):
        frac = torch.rand((E, raw.spec.action_dim), device='cuda', generator=g)
        o_r, r_r, _, _, _ = raw.step(frac * (hi - lo) + lo)
        o_w, r_w, _, _, _ = wrapped.step(frac)
        assert o_w.shape == (E, u._obs_dim) and r_w.shape == (E,)
        assert torch.equal(r_w, r_r.reshape(-1))
        # non-periodic columns: (x - min) / (max - min)
        j = 0
        ow = o_w.cpu().numpy(); orr = o_r.cpu().numpy()
        for col, (bi, n) in enumerate(u._out_entries):
            while names_raw[j] != n and not (n.endswith('_cos') or n.endswith('_sin')) and j < len(names_raw) - 1:
                j += 1
            if t['fn'][col] == S.OBS_FN_IDENTITY:
                exp = orr[:, j].astype('float64') * float(t['scale'][col]) + float(t['offset'][col])
                np.testing.assert_allclose(ow[:, col], exp, rtol=0, atol=2e-6)


def _reference_wrappers():
    """The reference's own `citylearn.wrappers` from oracle/_ref (installed by oracle/build_ref.py; travels to the GPU box)."""
    import sys
    ROOT = GOLDEN.parents[1]
    site = ROOT / 'oracle' / '_ref' / 'site'
    if not (site / 'citylearn' / 'wrappers.py').is_file():
        pytest.skip('oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)')
    for q in (ROOT / 'oracle' / 'shims', site):
        if str(q) not in sys.path:
            sys.path.insert(0, str(q))
    import citylearn.wrappers as RW
    return RW


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_reference_wrappers_run_unmodified_on_this_env(case):
    """SURVEY §8b "metadata facade": the UNMODIFIED reference wrappers (citylearn/wrappers.py:15-238) wrap this env at num_envs = 1 -
    they walk `env.unwrapped.buildings[i].observations(normalize=..., periodic_normalization=...)`, `estimate_observation_space(normalize=True)`
    and `action_space` - and return what they return around the reference env (tests/golden/wrappers)."""
    RW = _reference_wrappers()
    from citylearn_b200 import CityLearnEnv
    z, cfg = load(case)
    sch, src, ov = schema_for(cfg)
    base = CityLearnEnv(sch, data_source=src, num_envs=1, **ov)
    env = getattr(RW, cfg['wrapper'])(base)
    assert [n for row in env.observation_names for n in row] == [n for row in cfg['observation_names'] for n in row]
    assert np.array_equal(np.concatenate([s.low for s in env.observation_space]), z['space_low'])
    assert np.array_equal(np.concatenate([s.high for s in env.observation_space]), z['space_high'])
    obs, _ = env.reset()
    np.testing.assert_allclose(np.array([v for row in obs for v in row], dtype='float64'), z['reset_obs'], rtol=0, atol=2e-6)
    sizes = [len(r) for r in base.action_names]
    for k in range(len(z['actions'])):
        a = [float(x) for x in z['actions'][k]]
        nested, o = [], 0
        for n in sizes:
            nested.append(a[o:o + n])
            o += n
        obs, rew, term, _, _ = env.step(nested)
        np.testing.assert_allclose(np.array([v for row in obs for v in row], dtype='float64'), z['obs'][k], rtol=0, atol=2e-6)
        np.testing.assert_allclose(np.array(rew, dtype='float32'), z['reward'][k], rtol=1e-5, atol=1e-5)
