"""Where the time of one end-to-end host step goes (C2: 17 x 4096 envs): wall clock per call of the pieces `cl_step_host` chains.
Usage: python tools/e2e_breakdown.py [envs]"""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from citylearn_b200 import CityLearnEnv  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E)
A, L, R = env.spec.action_dim, env._obs_dim, env._reward_dim
pin = env.pinned_actions(1)[0]
pin[...] = np.random.RandomState(0).uniform(-1, 1, size=pin.shape)
page = pin.copy()
N = 300


def wall(fn, n=N):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def fresh():
    if env.time_step > env.time_steps - 400:
        env.reset()


st = torch.cuda.current_stream().cuda_stream
act_dev, rew_dev, out_pin = env._act, env._reward, env._out_pinned
print(f'envs {E}: H2D {E * A * 4} B, D2H {(E * R + L) * 4} B per step')
print('H2D pinned + sync       %7.1f us' % wall(lambda: (act_dev.copy_(torch.from_numpy(pin), non_blocking=True), torch.cuda.synchronize())))
print('D2H rewards+row + sync  %7.1f us' % wall(lambda: (env._out_pinned[E * L:].copy_(env._out[E * L:], non_blocking=True), torch.cuda.synchronize())))


def kernel_only():
    fresh()
    env._h.step(act_dev.data_ptr(), None, rew_dev.data_ptr(), env._district.data_ptr(), None, st)
    torch.cuda.synchronize()


env.reset()
print('cl_step (K = 1) + sync  %7.1f us' % wall(kernel_only))
env.reset()
print('step_host, pinned in    %7.1f us' % wall(lambda: (fresh(), env.step_host(pin))))
env.reset()
print('step_host, pageable in  %7.1f us' % wall(lambda: (fresh(), env.step_host(page))))
for mode, label in ((0, 'DMA both ways, sync'), (4, 'DMA both ways, flag'), (1, 'read in place'), (2, 'write in place'), (3, 'both in place'),
                    (6, 'write in place, flag'), (7, 'both in place, flag'), (5, 'read in place, flag')):
    env.reset()
    env._host_in_place = mode
    print('step_host mode %d %-22s %7.1f us' % (mode, label, wall(lambda: (fresh(), env.step_host(pin)))))
env._host_in_place = 3
env.reset()
print('step_host, full obs     %7.1f us' % wall(lambda: (fresh(), env.step_host(pin, full_observations=True))))
env.reset()
env._host_fast = False
print('step_host, torch-staged %7.1f us' % wall(lambda: (fresh(), env.step_host(page))))
