"""C4 (BASELINE.json configs[3]): synthetic 1024-building district, device-resident cl_rollout, per-GPU share of 8192 envs.

Sweeps the building-tile count (CTAs per cluster) through CL_B200_TILES; prints one JSON line per variant:
units/s, ms/step and the achieved fraction of the measured HBM peak for the algorithmic bytes of SURVEY.md §8d."""
import json
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from citylearn_b200 import CityLearnEnv, schema as S
from citylearn_b200.synthetic import make_wide_district

ROOT = Path(__file__).resolve().parents[1]


def run(spec, E, K, tiles, precision, with_obs=True):
    if tiles:
        os.environ['CL_B200_TILES'] = str(tiles)
    else:
        os.environ.pop('CL_B200_TILES', None)
    env = CityLearnEnv(spec, num_envs=E, precision=precision)
    geo = env._h.geometry()
    B, A, L = spec.n_buildings, spec.action_dim, env._obs_dim
    acts = torch.rand((K, E, A), device='cuda') * 2 - 1
    obs = torch.empty((K, E, L), device='cuda') if with_obs else None
    rew = torch.empty((K, E, env._reward_dim), device='cuda')
    dist = torch.empty((K, E, 3), device='cuda')
    best = 1e9
    W = 4
    for rep in range(3):
        env.reset()
        env.rollout(acts[:W].contiguous(), None if obs is None else obs[:W], rew[:W], dist[:W])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); env.rollout(acts[W:].contiguous(), None if obs is None else obs[W:], rew[W:], dist[W:]); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / (K - W))
    state = 20 if precision == 'fp64' else 12
    bpu = 4 * (A / B) + 2 * state + 4 * (L / B if with_obs else 0) + 4 + 12.0 / B
    peak = 6576.4
    try:
        peak = float(json.loads((ROOT / 'MEASURED_PEAKS.json').read_text()).get('hbm_gbs', peak))
    except Exception:
        pass
    gbs = bpu * B * E / (best * 1e-3) / 1e9
    print(json.dumps({'config': f'C4 synthetic {B} buildings x {E} envs', 'precision': precision, 'geometry': geo, 'obs': with_obs,
                      'ms_per_step': best, 'building_env_steps_per_s': B * E / best * 1e3, 'bytes_per_unit': bpu,
                      'hbm_gbs_algorithmic': gbs, 'hbm_frac': gbs / peak, 'checksum': float(rew.sum().item())}), flush=True)
    del env, acts, obs, rew, dist
    torch.cuda.empty_cache()


if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    tiles = [int(x) for x in sys.argv[4].split(',')] if len(sys.argv) > 4 else [0]
    sch, src = make_wide_district(N)
    spec = S.load(sch, data_source=src)
    for t in tiles:
        for precision in ('fp64', 'fp32'):
            try:
                run(spec, E, K, t, precision)
            except Exception as e:
                print(json.dumps({'tiles': t, 'precision': precision, 'error': repr(e)}), flush=True)
