"""A/B timing of kernel build variants (citylearn_b200/variants/*.so) on the C2 workload: one subprocess per (variant, precision)
so that each loads its own library (CL_B200_LIB).  Prints µs/step of a K-step cl_rollout, best of 3."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from citylearn_b200 import CityLearnEnv
E, K = 4096, 200
env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E, precision=sys.argv[1])
A, L = env.spec.action_dim, env._obs_dim
acts = torch.rand((K, E, A), device='cuda') * 2 - 1
obs = torch.empty((K, E, L), device='cuda'); rew = torch.empty((K, E, 17), device='cuda'); dist = torch.empty((K, E, 3), device='cuda')
best = 1e9
for rep in range(4):
    env.reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(acts, obs, rew, dist); e1.record(); torch.cuda.synchronize()
    if rep: best = min(best, e0.elapsed_time(e1) / K * 1e3)
print(json.dumps({'us_per_step': best, 'checksum': float(rew.sum().item())}))
''' % str(ROOT)

if __name__ == '__main__':
    libs = sorted((ROOT / 'citylearn_b200' / 'variants').glob('*.so'))
    for lib in libs:
        for precision in ('fp64', 'fp32'):
            env = dict(os.environ, CL_B200_LIB=str(lib))
            r = subprocess.run([sys.executable, '-c', CHILD, precision], capture_output=True, text=True, env=env)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
            print(lib.name, precision, line, flush=True)
