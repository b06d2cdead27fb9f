"""Phase timing inside the host-buffer step."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from citylearn_b200 import CityLearnEnv
E = 4096
env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E)
A = env.spec.action_dim
host = np.random.RandomState(0).uniform(-1, 1, size=(400, E, A)).astype('float32')
pc = time.perf_counter
def run(label, n=150, sync_between=False):
    env.reset(); torch.cuda.synchronize()
    T = np.zeros((n, 5))
    for i in range(n):
        t0 = pc()
        a, _ = env._parse_actions(host[i])
        t1 = pc()
        env._h.step(a.data_ptr(), env._obs.data_ptr(), env._reward.data_ptr(), env._district.data_ptr(), None, env._stream())
        env.time_step += 1
        t2 = pc()
        env._out_pinned.copy_(env._out, non_blocking=True)
        t3 = pc()
        torch.cuda.current_stream().synchronize()
        t4 = pc()
        x = float(env._reward_host[0, 0])
        t5 = pc()
        T[i] = [t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4]
    T *= 1e6
    print(label, 'mean us: parse %.1f  cl_step %.1f  d2h-issue %.1f  sync %.1f  read %.1f | total %.1f  | max total %.1f  p90 %.1f' % (
        *T[20:].mean(axis=0), T[20:].sum(axis=1).mean(), T[20:].sum(axis=1).max(), np.percentile(T[20:].sum(axis=1), 90)))
    return T
for rep in range(4):
    T = run(f'rep{rep}')
    tot = T.sum(axis=1); idx = np.where(tot > 2000)[0]; print('   outliers at', idx.tolist(), np.round(T[idx]).tolist())
print(np.round(T[20:40].sum(axis=1)))
t0 = pc()
for i in range(150): env.reset() if False else None
env.reset()
t0 = pc()
for i in range(150): env.step_host(host[i])
print('step_host loop mean us', (pc() - t0) / 150 * 1e6)
import os; print('cpus', os.cpu_count(), 'torch threads', torch.get_num_threads())
