"""C3 (BASELINE.json configs[2]): 2023 schema, 3 LSTM buildings x E envs, MARL; device-resident cl_rollout past the LSTM warm-up."""
import json
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from citylearn_b200 import CityLearnEnv
from citylearn_b200.data import DataSet

E = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 28
precision = sys.argv[3] if len(sys.argv) > 3 else 'fp64'
src = DataSet.get_source('citylearn_challenge_2023_phase_2_local_evaluation')
sch = src.schema()
sch['reward_function'] = {'type': 'citylearn.reward_function.MARL', 'attributes': {}}
env = CityLearnEnv(sch, data_source=src, central_agent=False, num_envs=E, precision=precision)
B, A, L = env.spec.n_buildings, env.spec.action_dim, env._obs_dim
lo = torch.tensor([v for b in env.spec.buildings for v in b.action_low], device='cuda')
hi = torch.tensor([v for b in env.spec.buildings for v in b.action_high], device='cuda')
acts = lo + torch.rand((K, E, A), device='cuda') * (hi - lo)
obs = torch.empty((K, E, L), device='cuda')
rew = torch.empty((K, E, env._reward_dim), device='cuda')
W = 16
best = 1e9
for rep in range(2):
    env.reset()
    env.rollout(acts[:W].contiguous(), obs[:W], rew[:W], None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); env.rollout(acts[W:].contiguous(), obs[W:], rew[W:], None); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / (K - W))
print(json.dumps({'config': f'C3 2023 LSTM {B} x {E}', 'precision': precision, 'ms_per_step': best, 'building_env_steps_per_s': B * E / best * 1e3,
                  'geometry': env._h.geometry()}))
