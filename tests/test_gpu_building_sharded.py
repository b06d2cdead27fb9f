"""GPU tests (need 2 devices): building-sharded districts - the district sums are completed across GPUs inside the step kernel
through peer memory (`cl_exchange_*`), checked against the single-GPU district (1e-5 of the district scale)."""
import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu
PALL = 'citylearn_challenge_2022_phase_all'
MARL = {'reward_function': 'citylearn.reward_function.MARL', 'reward_function_kwargs': None}


def _need_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 CUDA devices')


def _shards(E, world=2, **kw):
    from citylearn_b200.distributed import BuildingShardedEnv
    shards = [BuildingShardedEnv(PALL, E, device=f'cuda:{r}', rank=r, world=world, connect=False, **kw) for r in range(world)]
    BuildingShardedEnv.connect_in_process(shards)
    return shards


@pytest.mark.parametrize('reward', ['marl', 'default'])
def test_peer_memory_exchange_matches_single_gpu_district(reward):
    _need_two_gpus()
    from citylearn_b200 import CityLearnEnv
    import citylearn_b200.reward_function as rf
    E, K = 256, 30
    kw = {'reward_function': rf.MARL} if reward == 'marl' else {}
    full = CityLearnEnv(PALL, num_envs=E, device='cuda:0', central_agent=False, **kw)
    shards = _shards(E, **kw)
    B = full.spec.n_buildings
    assert sum(s.count for s in shards) == B and shards[0].first == 0 and shards[1].first == shards[0].count
    rng = np.random.RandomState(3)
    scale = 20.0                                             # district net of 17 buildings is O(20 kWh)
    for k in range(K):
        a = torch.from_numpy(rng.uniform(-1, 1, size=(E, B)).astype('float32'))
        _, rew, _, _, _ = full.step(a.to('cuda:0'))
        outs = [s.step(a[:, s.first:s.first + s.count].to(s.env.device)) for s in shards]     # async launches: rank 0 spins until rank 1 runs
        d_ref = full.district.cpu().numpy()
        for s, (_, r, _, _, _) in zip(shards, outs):
            d = s.district.cpu().numpy()
            assert np.max(np.abs(d - d_ref)) <= 1e-5 * scale, (k, s.rank)
            ref = rew[:, s.first:s.first + s.count].cpu().numpy()
            got = r.cpu().numpy()
            tol = 1e-5 * np.maximum(np.abs(ref), 1.0) if reward == 'marl' else 0.0
            assert np.all(np.abs(got - ref) <= tol), (k, s.rank, float(np.max(np.abs(got - ref))))
        # the exchanged sums are the same bits on every rank (rank-ordered addition)
        assert np.array_equal(shards[0].district.cpu().numpy(), shards[1].district.cpu().numpy())
    for s in shards:
        st = s.exchange_status()
        assert st['timeouts'] == 0 and st['steps_exchanged'] == K


def test_peer_memory_exchange_inside_a_rollout_launch():
    """K steps in ONE persistent launch per GPU: the kernels on the two devices meet at every step through peer memory."""
    _need_two_gpus()
    from citylearn_b200 import CityLearnEnv
    import citylearn_b200.reward_function as rf
    E, K = 512, 48
    full = CityLearnEnv(PALL, num_envs=E, device='cuda:0', central_agent=False, reward_function=rf.MARL)
    shards = _shards(E, reward_function=rf.MARL)
    B = full.spec.n_buildings
    acts = torch.rand((K, E, B)) * 2 - 1
    rew = torch.empty((K, E, B), device='cuda:0'); dst = torch.empty((K, E, 3), device='cuda:0')
    full.rollout(acts.to('cuda:0'), None, rew, dst)
    outs = []
    for s in shards:
        dev = s.env.device
        a = acts[:, :, s.first:s.first + s.count].contiguous().to(dev)
        r = torch.empty((K, E, s.count), device=dev); d = torch.empty((K, E, 3), device=dev)
        s.rollout(a, None, r, d)
        outs.append((r, d))
    for s, (r, d) in zip(shards, outs):
        assert float((d.cpu() - dst.cpu()).abs().max()) <= 2e-4
        ref = rew[:, :, s.first:s.first + s.count].cpu()
        assert float(((r.cpu() - ref).abs() / ref.abs().clamp_min(1.0)).max()) <= 1e-5
        assert s.exchange_status()['timeouts'] == 0
    assert torch.equal(outs[0][1].cpu(), outs[1][1].cpu())


def test_exchange_argument_checks():
    from citylearn_b200 import CityLearnEnv
    env = CityLearnEnv(PALL, num_envs=8, central_agent=True)
    with pytest.raises(NotImplementedError, match='central'):
        env._h.exchange_create(2, 0)
    env = CityLearnEnv(PALL, num_envs=8, central_agent=False)
    with pytest.raises(ValueError):
        env._h.exchange_create(1, 0)
    with pytest.raises(RuntimeError, match='cl_exchange_create'):
        env._h.exchange_connect(b'\0' * 128)
