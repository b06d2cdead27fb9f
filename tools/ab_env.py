"""A/B timing of run-time switches (environment variables read by cl_create) on the C2 workload: one subprocess per setting.
Usage: python tools/ab_env.py "NAME=VALUE,..." ...   (an empty string = defaults).  Prints us/step of a K-step cl_rollout, best of 3."""
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, json, torch
sys.path.insert(0, %r)
from citylearn_b200 import CityLearnEnv
E, K = int(sys.argv[2]), int(sys.argv[3])
env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E, precision=sys.argv[1])
A, L = env.spec.action_dim, env._obs_dim
g = torch.Generator(device='cuda').manual_seed(1)
acts = torch.rand((K, E, A), device='cuda', generator=g) * 2 - 1
obs = torch.empty((K, E, L), device='cuda'); rew = torch.empty((K, E, 17), device='cuda'); dist = torch.empty((K, E, 3), device='cuda')
best = 1e9
for rep in range(4):
    env.reset()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(acts, obs, rew, dist); e1.record(); torch.cuda.synchronize()
    if rep: best = min(best, e0.elapsed_time(e1) / K * 1e3)
print(json.dumps({'us_per_step': round(best, 3), 'geometry': env._h.geometry(), 'checksum': float(rew.double().sum().item()), 'obs_sum': float(obs.double().sum().item())}))
''' % str(ROOT)

if __name__ == '__main__':
    settings = sys.argv[1:] or ['']
    for E, K in ((4096, 200), (32768, 40)):
        for precision in ('fp64', 'fp32'):
            for s in settings:
                env = dict(os.environ)
                for kv in filter(None, s.split(',')):
                    k, v = kv.split('=', 1)
                    env[k] = v
                r = subprocess.run([sys.executable, '-c', CHILD, precision, str(E), str(K)], capture_output=True, text=True, env=env)
                line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]
                print(f'E={E} {precision} [{s or "default"}] {line}', flush=True)
