"""Break down the host-buffer step (env.step_host) on the GPU box."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np, torch
from citylearn_b200 import CityLearnEnv

E = 4096
env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E)
A = env.spec.action_dim
host = np.random.RandomState(0).uniform(-1, 1, size=(400, E, A)).astype('float32')
dev = torch.from_numpy(host[:50]).cuda()

def timeit(fn, n=100, warm=10):
    for i in range(warm): fn(i)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(warm + i)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e6

env.reset()
print('step(device tensor) + sync      us', timeit(lambda i: (env.step(dev[i % 50]), torch.cuda.synchronize())))
env.reset()
print('step(ndarray) + sync            us', timeit(lambda i: (env.step(host[i]), torch.cuda.synchronize())))
env.reset()
print('step_host                       us', timeit(lambda i: env.step_host(host[i])))
print('D2H obs only + sync             us', timeit(lambda i: (env._obs_pinned.copy_(env._obs, non_blocking=True), torch.cuda.synchronize())))
print('D2H reward only + sync          us', timeit(lambda i: (env._reward_pinned.copy_(env._reward, non_blocking=True), torch.cuda.synchronize())))
big = torch.empty((E, env._obs_dim)).pin_memory()
print('D2H obs into fresh pinned       us', timeit(lambda i: (big.copy_(env._obs, non_blocking=True), torch.cuda.synchronize())))
import ctypes
print('pinned? ', env._obs_pinned.is_pinned(), big.is_pinned())
print('D2H fused out (obs+reward) + sync us', timeit(lambda i: (env._out_pinned.copy_(env._out, non_blocking=True), torch.cuda.synchronize())))
n = env._out.numel()
p2 = torch.empty(n, dtype=torch.float32, pin_memory=True)
print('D2H fused into empty(pin_memory=True) us', timeit(lambda i: (p2.copy_(env._out, non_blocking=True), torch.cuda.synchronize())))
d2 = torch.empty(n, device='cuda')
print('D2H from fresh device buffer       us', timeit(lambda i: (p2.copy_(d2, non_blocking=True), torch.cuda.synchronize())))
env.reset()
def manual(i):
    env.step(host[i]); env._out_pinned.copy_(env._out, non_blocking=True); torch.cuda.current_stream().synchronize()
print('manual step + fused D2H + sync     us', timeit(manual))
def manual2(i):
    env.step(host[i]); env._obs_pinned.copy_(env._obs, non_blocking=True); env._reward_pinned.copy_(env._reward, non_blocking=True); torch.cuda.current_stream().synchronize()
env.reset()
print('manual step + 2 D2H + sync         us', timeit(manual2))
env.reset()
print('step_host again                    us', timeit(lambda i: env.step_host(host[i])))
