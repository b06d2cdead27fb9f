"""GPU tests: device-resident time step (`cl_advance_device`) and CUDA-graph closed loops (policy <-> env without the host)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

from helpers import max_abs_diff                            # noqa: E402

pytestmark = pytest.mark.gpu
PALL = 'citylearn_challenge_2022_phase_all'


def _policy(env, seed=0, dtype=None):
    from citylearn_b200.closed_loop import PerBuildingMLP
    B = env.spec.n_buildings
    return PerBuildingMLP(B, env._obs_dim // B, 1, hidden=64, dtype=dtype or torch.float32, device=env.device, seed=seed)


@pytest.mark.parametrize('stale', [True, False])
def test_graph_closed_loop_equals_host_driven_loop(stale):
    """K steps of [policy -> step] replayed from a CUDA graph give the rewards / observations / state of the same loop driven from
    the host through `env.step` (same kernels, same arithmetic: bit-identical), and the host time step follows the device counter."""
    from citylearn_b200 import CityLearnEnv
    from citylearn_b200.closed_loop import ClosedLoop
    E, K = 96, 21
    a = CityLearnEnv(PALL, num_envs=E, stale_observations=stale)
    b = CityLearnEnv(PALL, num_envs=E, stale_observations=stale)
    pol = _policy(a)
    loop = ClosedLoop(a, pol, steps_per_replay=4)
    ret = loop.run(K)                                   # 5 replays of 4 + 1 single step
    assert a.time_step == K and a._h.time_step() == K
    obs = b.observations
    acc = torch.zeros_like(b._reward)
    for _ in range(K):
        with torch.no_grad():
            act = pol(obs)
        obs, rew, _, _, _ = b.step(act)
        acc += rew
    assert max_abs_diff(ret.cpu().numpy(), acc.cpu().numpy()) == 0.0
    assert max_abs_diff(a.observations.cpu().numpy(), b.observations.cpu().numpy()) == 0.0
    assert torch.equal(a.state_dict()['state'], b.state_dict()['state'])
    # host-driven steps after graph replays continue from the device counter
    act = torch.zeros((E, a.spec.action_dim), device='cuda')
    a.step(act); b.step(act)
    assert a.time_step == K + 1 and a._h.time_step() == K + 1
    assert max_abs_diff(a._reward.cpu().numpy(), b._reward.cpu().numpy()) == 0.0


def test_advance_device_past_the_episode_end_is_a_no_op():
    from citylearn_b200 import CityLearnEnv
    E, T = 8, 6
    env = CityLearnEnv(PALL, num_envs=E, episode_time_steps=T)
    assert env.time_steps == T
    h = env._h
    st = torch.cuda.current_stream().cuda_stream
    h.device_time_enable(st)
    act = torch.zeros((E, env.spec.action_dim), device='cuda')
    rew = torch.zeros((E, env.spec.n_buildings), device='cuda')
    for k in range(T - 1):
        h.advance_device(1, act.data_ptr(), env._obs.data_ptr(), rew.data_ptr(), None, st)
    assert h.time_step() == T - 1
    last = rew.clone()
    state = env.state_dict()['state'].clone()
    rew.fill_(123.0)
    h.advance_device(1, act.data_ptr(), env._obs.data_ptr(), rew.data_ptr(), None, st)      # past the end: nothing written
    torch.cuda.synchronize()
    assert h.time_step() == T - 1
    assert float(rew.min()) == 123.0 and float(last.abs().sum()) > 0
    assert torch.equal(env.state_dict()['state'], state)
    with pytest.raises(RuntimeError):
        h.step(act.data_ptr(), env._obs.data_ptr(), rew.data_ptr(), None, None, st)       # the host path still refuses
    env.time_step = T - 1
    env.reset()
    assert h.time_step() == 0


def test_advance_device_needs_enable():
    from citylearn_b200 import CityLearnEnv
    env = CityLearnEnv(PALL, num_envs=4)
    act = torch.zeros((4, env.spec.action_dim), device='cuda')
    with pytest.raises(RuntimeError, match='cl_device_time_enable'):
        env._h.advance_device(1, act.data_ptr(), None, None, None, torch.cuda.current_stream().cuda_stream)


def test_auto_reset_and_mask_semantics():
    """Batched episode end: opt-in same-step auto-reset; a partial reset mask is refused (lock-step envs)."""
    from citylearn_b200 import CityLearnEnv
    E, T = 16, 5
    env = CityLearnEnv(PALL, num_envs=E, episode_time_steps=T, auto_reset=True)
    ref = CityLearnEnv(PALL, num_envs=E, episode_time_steps=T)
    act = torch.zeros((E, env.spec.action_dim), device='cuda')
    for k in range(T - 1):
        obs, rew, term, _, info = env.step(act)
        o2, r2, t2, _, _ = ref.step(act)
        assert term == t2 and torch.equal(rew, r2)
    assert term and env.time_step == 0 and torch.equal(info['final_observation'], o2)
    ref.reset()                                                      # the reference-style caller resets itself: same next episode
    assert torch.equal(obs, ref.observations) and env.episode_tracker.episode == ref.episode_tracker.episode
    obs, rew, term, _, _ = env.step(act)                             # and the next episode runs
    assert not term and env.time_step == 1
    with pytest.raises(NotImplementedError, match='lock-step'):
        env.reset(mask=torch.arange(E) < 3)
    env.reset(mask=torch.ones(E, dtype=torch.bool))
    with pytest.raises(ValueError):
        CityLearnEnv(PALL, num_envs=1, auto_reset=True)


def test_devices_keyword_shards_envs_over_the_devices_of_one_process():
    from citylearn_b200 import CityLearnEnv
    from citylearn_b200.distributed import DeviceShardedEnv
    one = CityLearnEnv(PALL, num_envs=8, devices=['cuda:0'])
    assert isinstance(one, CityLearnEnv) and one.device == torch.device('cuda', 0)
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')
    E = 64
    fleet = CityLearnEnv(PALL, num_envs=E, devices=['cuda:0', 'cuda:1'])
    assert isinstance(fleet, DeviceShardedEnv) and [e.num_envs for e in fleet.envs] == [32, 32]
    ref = CityLearnEnv(PALL, num_envs=E, device='cuda:0')
    g = torch.Generator().manual_seed(0)
    for k in range(6):
        a = torch.rand((E, ref.spec.action_dim), generator=g) * 2 - 1
        obs, rew, term, _, _ = fleet.step(a)
        o2, r2, t2, _, _ = ref.step(a.cuda())
        assert obs[1].device == torch.device('cuda', 1)
        assert torch.equal(DeviceShardedEnv.gather(rew, 'cuda:0'), r2) and torch.equal(DeviceShardedEnv.gather(obs, 'cuda:0'), o2)
    assert fleet.time_step == 6 and fleet.observation_names == ref.observation_names
