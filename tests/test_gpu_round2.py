"""GPU tests of the round-2 paths (through the C ABI): distinct-action oracle parity at full BASELINE sizes, the 1024-thread
instantiation, the fresh-observation table path, the shared-row host paths, checkpoints across episode windows, argument checks."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from citylearn_b200.schema import DYN                       # noqa: E402
from helpers import max_abs_diff                            # noqa: E402

pytestmark = pytest.mark.gpu
PALL = 'citylearn_challenge_2022_phase_all'


def _c3_schema():
    from citylearn_b200.data import DataSet
    src = DataSet.get_source('citylearn_challenge_2023_phase_2_local_evaluation')
    sch = src.schema()
    sch['reward_function'] = {'type': 'citylearn.reward_function.MARL', 'attributes': {}}
    return sch, src


@pytest.mark.parametrize('E,threads', [(4096, 512), (16384, 992)])
def test_c2_distinct_actions_match_oracle_at_full_size(E, threads):
    """BASELINE configs[1] (17 x 4096) and the 1024-thread / 64-register instantiation `cl_create` selects beyond two waves
    (17 x 16384): every env has its own action sequence; rewards, district sums and the physics trace equal the oracle's bit for bit."""
    from citylearn_b200 import CityLearnEnv
    from citylearn_oracle import OracleEnv
    env = CityLearnEnv(PALL, num_envs=E, debug_trace=True)
    assert env._h.geometry()['threads'] == threads, env._h.geometry()
    oracle = OracleEnv(env.spec, E)
    assert max_abs_diff(env.reset()[0].cpu().numpy(), oracle.reset().astype('float32')) == 0.0
    rng = np.random.RandomState(31)
    for k in range(12):
        a = rng.uniform(-1, 1, size=(E, env.spec.action_dim)).astype('float32')
        if k == 5:
            a[::3] = 0.0                            # idle batteries: zero numerators everywhere
        obs, rew, _, _, _ = env.step(torch.from_numpy(a).cuda())
        oobs, orew, odist, odyn = oracle.step(a)
        assert max_abs_diff(obs.cpu().numpy(), oobs) == 0.0
        assert np.array_equal(rew.cpu().numpy(), orew), k
        assert np.array_equal(env.district.cpu().numpy(), odist), k
        tr = env.trace.cpu().numpy()
        for n in ('electrical_storage_soc', 'electrical_storage_energy_balance', 'net_electricity_consumption',
                  'electrical_storage_degraded_capacity'):
            assert np.array_equal(tr[..., DYN[n]], odyn[..., DYN[n]].astype('float32')), (n, k)


def test_c3_distinct_actions_match_oracle_at_2048_envs():
    """BASELINE configs[2] shape (3 LSTM buildings, MARL) at 2048 envs, past the LSTM warm-up."""
    from citylearn_b200 import CityLearnEnv
    from citylearn_oracle import OracleEnv
    sch, src = _c3_schema()
    E, K = 2048, 20
    env = CityLearnEnv(sch, data_source=src, central_agent=False, num_envs=E, debug_trace=True)
    oracle = OracleEnv(env.spec, E)
    assert max_abs_diff(env.reset()[0].cpu().numpy(), oracle.reset().astype('float32')) == 0.0
    rng = np.random.RandomState(41)
    lo = np.concatenate([b.action_low for b in env.spec.buildings])
    hi = np.concatenate([b.action_high for b in env.spec.buildings])
    for k in range(K):
        a = (lo + rng.uniform(0, 1, size=(E, env.spec.action_dim)) * (hi - lo)).astype('float32')
        obs, rew, _, _, _ = env.step(a)
        oobs, orew, odist, odyn = oracle.step(a)
        assert max_abs_diff(obs.cpu().numpy(), oobs) == 0.0
        tr = env.trace.cpu().numpy()
        for n in ('electrical_storage_soc', 'dhw_storage_soc', 'net_electricity_consumption', 'cooling_electricity_consumption', 'cooling_demand'):
            assert np.array_equal(tr[..., DYN[n]], odyn[..., DYN[n]].astype('float32')), (n, k)
        assert max_abs_diff(tr[..., DYN['indoor_dry_bulb_temperature']], odyn[..., DYN['indoor_dry_bulb_temperature']]) < 3e-5
        assert np.array_equal(rew.cpu().numpy(), orew)


@pytest.mark.parametrize('dataset,E,kw', [(PALL, 4096, {}), (PALL, 640, {'central_agent': True}),
                                          ('citylearn_challenge_2020_climate_zone_1', 256, {})])
def test_fresh_observations_table_path_matches_oracle(dataset, E, kw):
    """stale_observations=False through the per-env row images (TMA load of the table row, DYN columns patched in shared memory, one
    TMA store per env row): observations incl. soc / net of the step, rewards and district sums equal the oracle's, bit for bit."""
    from citylearn_b200 import CityLearnEnv
    from citylearn_oracle import OracleEnv
    env = CityLearnEnv(dataset, num_envs=E, stale_observations=False, **kw)
    oracle = OracleEnv(env.spec, E, stale_observations=False)
    assert max_abs_diff(env.reset()[0].cpu().numpy(), oracle.reset().astype('float32')) == 0.0
    rng = np.random.RandomState(5)
    lo = np.concatenate([b.action_low for b in env.spec.buildings])
    hi = np.concatenate([b.action_high for b in env.spec.buildings])
    K = 14
    acts = (lo + rng.uniform(0, 1, size=(K, E, env.spec.action_dim)) * (hi - lo)).astype('float32')
    ref = []
    for k in range(K):
        obs, rew, _, _, _ = env.step(torch.from_numpy(acts[k]).cuda())
        oobs, orew, odist, _ = oracle.step(acts[k])
        ref.append((oobs, orew))
        assert max_abs_diff(obs.cpu().numpy(), oobs) == 0.0, k
        assert np.array_equal(rew.cpu().numpy(), orew), k
        assert np.array_equal(env.district.cpu().numpy(), odist), k
    names = [n for _, n in env._entries]
    assert float(obs[:, names.index('electrical_storage_soc')].abs().sum()) > 0.0
    # the same through ONE rollout launch (double-buffered images across steps)
    env.reset()
    L, R = env._obs_dim, env._reward_dim
    o = torch.empty((K, E, L), device='cuda'); r = torch.empty((K, E, R), device='cuda')
    env.rollout(torch.from_numpy(acts).cuda(), o, r, None)
    for k in range(K):
        assert max_abs_diff(o[k].cpu().numpy(), ref[k][0]) == 0.0, k
        assert np.array_equal(r[k].cpu().numpy(), ref[k][1]), k


def test_fresh_observations_with_normalized_wrapper_table_vs_gather(monkeypatch):
    """Fused observation transforms on the patched DYN columns: table path == general gather path."""
    from citylearn_b200 import CityLearnEnv, wrappers as W
    E, K = 64, 12
    acts = np.random.RandomState(3).uniform(0, 1, size=(K, E, 5)).astype('float32')

    def run():
        env = W.NormalizedSpaceWrapper(CityLearnEnv('citylearn_challenge_2022_phase_1', num_envs=E, stale_observations=False))
        env.reset()
        return [env.step(torch.from_numpy(acts[k]).cuda())[0].clone() for k in range(K)]
    a = run()
    monkeypatch.setenv('CL_B200_OBS_TABLE_MB', '0')
    b = run()
    for k in range(K):
        assert torch.equal(a[k], b[k]), k


def test_step_host_shared_row_and_rollout_host():
    from citylearn_b200 import CityLearnEnv
    E, K = 96, 10
    acts = np.random.RandomState(8).uniform(-1, 1, size=(K, E, 17)).astype('float32')
    full = CityLearnEnv(PALL, num_envs=E)
    shared = CityLearnEnv(PALL, num_envs=E)
    block = CityLearnEnv(PALL, num_envs=E)
    assert shared.shared_observation_row
    full.reset(); shared.reset(); block.reset()
    obs_rows, rew_all, term = block.rollout_host(acts)
    assert obs_rows.shape == (K, full._obs_dim) and rew_all.shape == (K, E, 17) and not term
    for k in range(K):
        o1, r1, _ = full.step_host(acts[k], full_observations=True)
        o2, r2, _ = shared.step_host(acts[k])
        assert o2.shape == o1.shape and np.array_equal(o1, o2)
        assert np.array_equal(r1, r2)
        assert np.array_equal(obs_rows[k], o1[0]) and np.array_equal(rew_all[k], r1)
    assert torch.equal(shared.observations, full.observations) and torch.equal(block.observations, full.observations)
    assert block.time_step == K
    # per-env windows: no shared row
    env = CityLearnEnv('citylearn_challenge_2022_phase_1', num_envs=4, simulation_end_time_step=999, episode_time_steps=200)
    env.reset(options={'episode_start': torch.tensor([0, 100, 200, 300])})
    assert not env.shared_observation_row
    with pytest.raises(ValueError):
        env.step_host(np.zeros((4, 5), dtype='float32'), full_observations=False)
    o, _, _ = env.step_host(np.zeros((4, 5), dtype='float32'))
    assert o.shape == (4, env._obs_dim) and not np.array_equal(o[0], o[1])


def test_checkpoint_restores_episode_window_and_outage():
    """A checkpoint loaded into a FRESH env (other episode window after its own reset) continues on the checkpoint's rows."""
    from citylearn_b200 import CityLearnEnv
    kw = dict(num_envs=8, simulation_start_time_step=0, simulation_end_time_step=1999, episode_time_steps=400)
    a = CityLearnEnv('citylearn_challenge_2022_phase_1', **kw)
    a.reset(); a.reset()                            # second window: rows 400 .. 799
    assert a.episode_tracker.episode_start_time_step == 400
    acts = np.random.RandomState(1).uniform(-1, 1, size=(30, 8, 5)).astype('float32')
    for k in range(10):
        a.step(acts[k])
    sd = a.state_dict()
    ref = [tuple(x.clone() for x in a.step(acts[k])[:2]) for k in range(10, 30)]
    b = CityLearnEnv('citylearn_challenge_2022_phase_1', **kw)
    b.reset()                                       # window 0 .. 399
    b.load_state_dict(sd)
    assert b.time_step == 10 and b.episode_tracker.episode_start_time_step == 400
    for k in range(10, 30):
        o, r, _, _, _ = b.step(acts[k])
        assert torch.equal(o, ref[k - 10][0]) and torch.equal(r, ref[k - 10][1]), k
    with pytest.raises(RuntimeError):
        b.evaluate_batched()
    # per-env windows survive a checkpoint too
    c = CityLearnEnv('citylearn_challenge_2022_phase_1', **kw)
    starts = torch.tensor([0, 50, 100, 150, 200, 250, 300, 350])
    c.reset(options={'episode_start': starts})
    for k in range(5):
        c.step(acts[k])
    sd = c.state_dict()
    ref = [c.step(acts[k])[0].clone() for k in range(5, 12)]
    d = CityLearnEnv('citylearn_challenge_2022_phase_1', **kw)
    d.load_state_dict(sd)
    for k in range(5, 12):
        assert torch.equal(d.step(acts[k])[0], ref[k - 5])


def test_episode_start_validation():
    from citylearn_b200 import CityLearnEnv
    env = CityLearnEnv('citylearn_challenge_2022_phase_1', num_envs=4, simulation_end_time_step=999, episode_time_steps=200)
    for bad in ([0, 100, 200, 900], [-1, 0, 0, 0], [0, 1, 2]):
        with pytest.raises(ValueError):
            env.reset(options={'episode_start': torch.tensor(bad)})
    env.reset(options={'episode_start': torch.tensor([0, 100, 200, 800])})
    from citylearn_b200.data import DataSet
    src = DataSet.get_source('citylearn_challenge_2023_phase_2_local_evaluation')
    out = CityLearnEnv(src.schema(), data_source=src, num_envs=2)
    with pytest.raises(ValueError):
        out.reset(options={'episode_start': torch.tensor([0, 0])})


def test_step_host_single_call_path_equals_the_staged_path():
    """`cl_step_host` (one native call per step; in place on page-locked memory, or with DMA copies) against the torch-staged path (an
    env that tracks episode rewards keeps it): same observations / rewards from page-locked, pageable, non-contiguous and wrong-dtype
    actions; both observation modes."""
    from citylearn_b200 import CityLearnEnv
    E, K = 64, 12
    fast = CityLearnEnv(PALL, num_envs=E)
    copies = CityLearnEnv(PALL, num_envs=E)
    copies._host_in_place = 0                        # DMA copies before / after the kernel instead of in-place PCIe access
    slow = CityLearnEnv(PALL, num_envs=E, track_episode_rewards=True)
    assert fast._host_fast and fast._host_in_place == 3 and copies._host_fast and not slow._host_fast
    mixed = CityLearnEnv(PALL, num_envs=E)
    mixed._host_in_place = 6                         # DMA in, in-place out, polled completion
    mp = mixed.pinned_actions(1)[0]
    cp = copies.pinned_actions(1)[0]
    rng = np.random.RandomState(4)
    pinned = fast.pinned_actions(2)
    n0 = fast.gpu_launches
    for k in range(K):
        a = rng.uniform(-1, 1, size=(E, 17)).astype('float32')
        if k % 4 == 0:
            np.copyto(pinned[k % 2], a); arg = pinned[k % 2]
        elif k % 4 == 1:
            arg = a
        elif k % 4 == 2:
            arg = np.asfortranarray(a)                   # not C-contiguous: staged
        else:
            arg = a.astype('float64')                    # wrong dtype: staged
        full = k >= K // 2
        o1, r1, t1 = fast.step_host(arg, full_observations=full or None)
        o2, r2, t2 = slow.step_host(a, full_observations=full or None)
        np.copyto(cp, a)
        o3, r3, t3 = copies.step_host(cp, full_observations=full or None)
        assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and t1 == t2, k
        assert np.array_equal(o3, o2) and np.array_equal(r3, r2) and t3 == t2, k
        np.copyto(mp, a)
        o4, r4, t4 = mixed.step_host(mp, full_observations=full or None)
        assert np.array_equal(o4, o2) and np.array_equal(r4, r2) and t4 == t2, k
    assert fast.time_step == slow.time_step == K and fast._h.time_step() == K
    assert fast.gpu_launches - n0 >= K
    assert torch.equal(fast.observations, slow.observations)
    # the device-side step continues from the same state
    act = torch.zeros((E, 17), device='cuda')
    assert torch.equal(fast.step(act)[1], slow.step(act)[1])
    with pytest.raises(ValueError):
        fast.step_host(np.zeros((E, 3), dtype='float32'))
    # episode end: the last step reports terminated, the next call raises like step()
    short = CityLearnEnv(PALL, num_envs=4, episode_time_steps=4)
    for k in range(3):
        _, _, term = short.step_host(np.zeros((4, 17), dtype='float32'))
    assert term
    with pytest.raises(RuntimeError):
        short.step_host(np.zeros((4, 17), dtype='float32'))


def test_split_phase_barrier_instantiation_is_bit_identical(monkeypatch):
    """`CL_B200_DECOUPLE=1` selects the barrier-free step (arrive at step k, wait one step later, lagged district sums): an opt-in
    experiment that must still give the same bits - observations, rewards, district sums, state - per step and inside one launch."""
    from citylearn_b200 import CityLearnEnv
    E, K = 300, 41                                           # several blocks, the last one partly filled
    ref = CityLearnEnv(PALL, num_envs=E)
    monkeypatch.setenv('CL_B200_DECOUPLE', '1')
    dec = CityLearnEnv(PALL, num_envs=E)
    dec2 = CityLearnEnv(PALL, num_envs=E)
    monkeypatch.delenv('CL_B200_DECOUPLE')
    g = torch.Generator(device='cuda').manual_seed(5)
    acts = torch.rand((K, E, 17), device='cuda', generator=g) * 2 - 1
    L = ref._obs_dim
    bufs = [(torch.zeros((K, E, L), device='cuda'), torch.zeros((K, E, 17), device='cuda'), torch.zeros((K, E, 3), device='cuda')) for _ in range(2)]
    ref.rollout(acts, *bufs[0])
    dec.rollout(acts, *bufs[1])
    for a, b in zip(bufs[0], bufs[1]):
        assert torch.equal(a, b)
    assert torch.equal(ref.state_dict()['state'], dec.state_dict()['state'])
    for k in range(6):                                       # single-step launches (K = 1) and short blocks
        o, r, _, _, _ = dec2.step(acts[k])
        assert torch.equal(o, bufs[0][0][k]) and torch.equal(r, bufs[0][1][k]) and torch.equal(dec2.district, bufs[0][2][k])
    o3 = torch.zeros((3, E, L), device='cuda'); r3 = torch.zeros((3, E, 17), device='cuda'); d3 = torch.zeros((3, E, 3), device='cuda')
    dec2.rollout(acts[6:9].contiguous(), o3, r3, d3)
    assert torch.equal(o3, bufs[0][0][6:9]) and torch.equal(r3, bufs[0][1][6:9]) and torch.equal(d3, bufs[0][2][6:9])


def test_lstm_tensor_core_cell_matches_the_scalar_cell(monkeypatch):
    """The mma.sync (3xTF32) LSTM cell against the scalar float32 cell (`CL_B200_NO_LSTM_MMA`): 40 steps, 28 of them with the LSTM
    live and feeding back its own predictions - indoor temperature within 4e-5 degC of each other (measured 2.1e-5), the energy path identical."""
    from citylearn_b200 import CityLearnEnv
    sch, src = _c3_schema()
    E, K = 1120, 40                                      # 7 blocks of 160 envs: whole warps on one building
    mma = CityLearnEnv(sch, data_source=src, central_agent=False, num_envs=E, debug_trace=True)
    monkeypatch.setenv('CL_B200_NO_LSTM_MMA', '1')
    ref = CityLearnEnv(sch, data_source=src, central_agent=False, num_envs=E, debug_trace=True)
    monkeypatch.delenv('CL_B200_NO_LSTM_MMA')
    g = torch.Generator(device='cuda').manual_seed(3)
    lo = torch.tensor(np.concatenate([b.action_low for b in mma.spec.buildings]), device='cuda')
    hi = torch.tensor(np.concatenate([b.action_high for b in mma.spec.buildings]), device='cuda')
    worst = 0.0
    for k in range(K):
        a = lo + torch.rand((E, mma.spec.action_dim), device='cuda', generator=g) * (hi - lo)
        o1, r1, _, _, _ = mma.step(a)
        o2, r2, _, _, _ = ref.step(a)
        t1, t2 = mma.trace, ref.trace
        worst = max(worst, float((t1[..., DYN['indoor_dry_bulb_temperature']] - t2[..., DYN['indoor_dry_bulb_temperature']]).abs().max()))
        for n in ('electrical_storage_soc', 'dhw_storage_soc', 'net_electricity_consumption', 'cooling_electricity_consumption', 'cooling_demand'):
            assert torch.equal(t1[..., DYN[n]], t2[..., DYN[n]]), (n, k)
        assert torch.equal(o1, o2) and torch.equal(r1, r2), k
    assert 0.0 < worst < 4e-5, worst                     # two float32 evaluation orders, each within 3e-5 of the reference (0.0: the tensor-core path did not run)


def test_multi_building_reward_function_matches_the_fused_rewards():
    """Per-building reward functions from the schema (`MultiBuildingRewardFunction`, citylearn/reward_function.py:90-117; Python path
    over the per-unit trace): every building's column equals the column of an env whose fused reward is that building's function."""
    from citylearn_b200 import CityLearnEnv
    from citylearn_b200 import reward_function as rf
    from citylearn_b200.data import DataSet
    src = DataSet.get_source('citylearn_challenge_2022_phase_1')
    sch = src.schema()
    names = list(sch['buildings'])
    sch['reward_function'] = {'type': {'default': 'citylearn.reward_function.RewardFunction', names[1]: 'citylearn.reward_function.SolarPenaltyReward',
                                       names[3]: 'citylearn.reward_function.IndependentSACReward'},
                              'attributes': {}}
    E = 24
    multi = CityLearnEnv(sch, data_source=src, num_envs=E)
    assert isinstance(multi.reward_function, rf.MultiBuildingRewardFunction) and multi._reward_id == -1
    plain = {k: CityLearnEnv('citylearn_challenge_2022_phase_1', num_envs=E, reward_function=getattr(rf, k))
             for k in ('RewardFunction', 'SolarPenaltyReward', 'IndependentSACReward')}
    which = ['RewardFunction'] * len(names)
    which[1], which[3] = 'SolarPenaltyReward', 'IndependentSACReward'
    g = torch.Generator(device='cuda').manual_seed(12)
    for k in range(15):
        a = torch.rand((E, multi.spec.action_dim), device='cuda', generator=g) * 2 - 1
        _, r, _, _, _ = multi.step(a)
        cols = {n: e.step(a)[1] for n, e in plain.items()}
        for b, n in enumerate(which):
            assert max_abs_diff(r[:, b].cpu().numpy(), cols[n][:, b].cpu().numpy()) <= 2e-6 * max(1.0, float(cols[n][:, b].abs().max())), (k, b, n)
