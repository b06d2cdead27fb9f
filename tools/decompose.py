"""Where does the per-step time go?  Times cl_rollout with outputs switched off, and at several env counts."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from citylearn_b200 import CityLearnEnv

def run(E, precision, with_obs, with_rew, with_dist, K=200):
    env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E, precision=precision)
    A, L, B = env.spec.action_dim, env._obs_dim, env.spec.n_buildings
    acts = torch.rand((K, E, A), device='cuda') * 2 - 1
    obs = torch.empty((K, E, L), device='cuda') if with_obs else None
    rew = torch.empty((K, E, B), device='cuda') if with_rew else None
    dst = torch.empty((K, E, 3), device='cuda') if with_dist else None
    best = 1e9
    for rep in range(3):
        env.reset()
        env.rollout(acts[:20].contiguous(), None if obs is None else obs[:20], None if rew is None else rew[:20], None if dst is None else dst[:20])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(acts, obs, rew, dst); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / K)
    return best

if __name__ == '__main__':
    for precision in ('fp32', 'fp64'):
        for E in (4096, 16384, 65536):
            K = 200 if E <= 16384 else 40
            full = run(E, precision, True, True, True, K)
            noobs = run(E, precision, False, True, True, K)
            none = run(E, precision, False, False, False, K)
            print(f'{precision} E={E}: full {full:.2f} us/step, no-obs {noobs:.2f}, physics-only {none:.2f}  | full units/s {17*E/full*1e6/1e9:.2f} G', flush=True)
