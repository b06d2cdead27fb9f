"""Throughput of the other BASELINE.json configs (not the headline bench line): C3 = 2023 schema (3 LSTM buildings) x 65536 envs,
C5-sized = 2022_phase_all x 32768 envs.  Device-resident cl_rollout, CUDA events."""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from citylearn_b200 import CityLearnEnv
from citylearn_b200.data import DataSet

def run(name, make, E, K, with_obs=True):
    env = make(E)
    B, A, L = env.spec.n_buildings, env.spec.action_dim, env._obs_dim
    lo = torch.tensor([v for b in env.spec.buildings for v in b.action_low], device='cuda')
    hi = torch.tensor([v for b in env.spec.buildings for v in b.action_high], device='cuda')
    acts = lo + torch.rand((K, E, A), device='cuda') * (hi - lo)
    obs = torch.empty((K, E, L), device='cuda') if with_obs else None
    rew = torch.empty((K, E, env._reward_dim), device='cuda')
    best = 1e9
    for rep in range(2):
        env.reset()
        W = 16                                     # past the LSTM warm-up (predictions start at t = 12)
        env.rollout(acts[:W].contiguous(), None if obs is None else obs[:W], rew[:W], None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(acts[W:].contiguous(), None if obs is None else obs[W:], rew[W:], None); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (K - W))
    out = {'config': name, 'buildings': B, 'envs': E, 'ms_per_step': best, 'building_env_steps_per_s': B * E / best * 1e3, 'precision': env.precision}
    print(json.dumps(out), flush=True)

def c3(E, precision='fp64'):
    src = DataSet.get_source('citylearn_challenge_2023_phase_2_local_evaluation')
    sch = src.schema(); sch['reward_function'] = {'type': 'citylearn.reward_function.MARL', 'attributes': {}}
    return CityLearnEnv(sch, data_source=src, central_agent=False, num_envs=E, precision=precision)

if __name__ == '__main__':
    run('C3 2023 LSTM x 8192 envs', lambda E: c3(E), 8192, 48)
    run('C3 2023 LSTM x 65536 envs', lambda E: c3(E), 65536, 40)
    run('C3 2023 LSTM x 65536 envs fp32', lambda E: c3(E, 'fp32'), 65536, 40)
    run('C5-size 2022_phase_all x 32768 envs', lambda E: CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E), 32768, 64)
