"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes (the N > 1 data path itself has no collective)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from citylearn_b200.distributed import fleet_reward_stats, shard_range


def test_shard_range_partitions_exactly():
    for total in (1, 2, 7, 4096, 4097, 32768):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (o0, c0), (o1, c1) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from citylearn_b200 import schema as S
        from citylearn_oracle import OracleEnv   # CPU stand-in for the per-rank shard (the product itself needs CUDA)
        spec = S.load('citylearn_challenge_2022_phase_1')
        offset, count = shard_range(total_envs, rank, world)
        env = OracleEnv(spec, count)
        env.reset()
        rng = np.random.RandomState(0)
        acts = rng.uniform(-1, 1, size=(12, total_envs, spec.action_dim)).astype('float32')[:, offset:offset + count]
        rsum = rmin = rmax = None
        for k in range(12):
            _, rew, _, _ = env.step(acts[k])
            r = torch.from_numpy(rew.astype('float32'))
            rsum = r.clone() if rsum is None else rsum + r
            rmin = r.clone() if rmin is None else torch.minimum(rmin, r)
            rmax = r.clone() if rmax is None else torch.maximum(rmax, r)
        stats = fleet_reward_stats(rsum, rmin, rmax, steps=12)
        if rank == 0:
            out.put({k: (v.numpy() if torch.is_tensor(v) else v) for k, v in stats.items()})
    finally:
        dist.destroy_process_group()


def test_env_sharded_fleet_matches_single_process():
    """2 gloo ranks, 5 envs split 3 + 2: fleet statistics equal the unsharded run (envs are independent)."""
    from citylearn_b200 import schema as S
    from citylearn_oracle import OracleEnv
    total = 5
    ctx = mp.get_context('spawn')
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    spec = S.load('citylearn_challenge_2022_phase_1')
    env = OracleEnv(spec, total)
    env.reset()
    rng = np.random.RandomState(0)
    acts = rng.uniform(-1, 1, size=(12, total, spec.action_dim)).astype('float32')
    rs = []
    for k in range(12):
        _, rew, _, _ = env.step(acts[k])
        rs.append(rew.astype('float32'))
    rs = np.stack(rs)                       # [K, E, B]
    assert got['n_envs'] == total
    np.testing.assert_allclose(got['sum_per_env_mean'], rs.sum(axis=0).astype('float64').mean(axis=0), rtol=1e-6)
    np.testing.assert_allclose(got['min'], rs.min(axis=(0, 1)), rtol=0)
    np.testing.assert_allclose(got['max'], rs.max(axis=(0, 1)), rtol=0)
