"""Cycle stamps inside advance_kernel (debug build with -DCL_PHASE_TIMING): where does one warp spend a step?"""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from citylearn_b200 import build as B
lib = ROOT / 'gpurun_tmp_timing.so'
if not lib.exists() or '--build' in sys.argv:
    cmd = [B.nvcc_path(), *B.NVCC_FLAGS, '-DCL_PHASE_TIMING', '-o', str(lib), str(B.SRC)]
    subprocess.run(cmd, check=True)
if '--build' in sys.argv:
    sys.exit(0)
os.environ['CL_B200_LIB'] = str(lib)
import torch
from citylearn_b200 import CityLearnEnv
for precision in ('fp64', 'fp32'):
    E, K = 4096, 64
    env = CityLearnEnv('citylearn_challenge_2022_phase_all', num_envs=E, precision=precision)
    acts = torch.rand((K, E, 17), device='cuda') * 2 - 1
    obs = torch.empty((K, E, env._obs_dim), device='cuda'); rew = torch.empty((K, E, 17), device='cuda')
    stamps = torch.zeros((K, E, 3), device='cuda')
    env.reset(); env.rollout(acts, obs, rew, stamps); env.reset(); stamps.zero_(); env.rollout(acts, obs, rew, stamps)
    torch.cuda.synchronize()
    st = stamps.flatten()[:K * 8].reshape(K, 8).cpu().double().numpy()
    names = ['top', 'rows ready', 'inputs+actions', 'physics done', 'red+tmpl done', 'after S1', 'district/reward done']
    import numpy as np
    d = np.diff(st[:, :7], axis=1)
    step = np.diff(st[:, 0])
    print(precision, 'cycles per step (median)', np.median(step), '=', np.median(step) / 1.965e3, 'us')
    for i in range(6):
        print(f'   {names[i]:>22s} -> {names[i+1]:<24s} median {np.median(d[8:, i]):8.0f} cycles')
    print(f'   {"district/reward done":>22s} -> {"next top (obs write)":<24s} median {np.median(st[9:, 0] - st[8:-1, 6]):8.0f} cycles')
