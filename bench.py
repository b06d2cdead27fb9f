#!/usr/bin/env python
"""bench.py - throughput of the CityLearn step path on B200 (driver contract: one JSON line on rank 0).

Workload (BASELINE.json configs[1]): citylearn_challenge_2022_phase_all, 17 buildings x 4096 parallel envs PER GPU
(weak scaling: envs shard across GPUs with no data-path collective), synthetic uniform(-1, 1) actions.
A "step" is one environment time step of all 17 x 4096 units of a rank: actions in, state update, district sums, reward,
observation at t+1 out.  Metric: building-env steps / s (whole job, all ranks).

  value     device-resident: K steps enqueued by ONE cl_rollout call (actions [K,E,A] already in HBM, every step writes
            its own observation / reward slab, so the K * 7.8 MB output stream is larger than the 126 MB L2),
            timed with CUDA events on the launch stream, max over ranks.
  e2e       the public API with HOST buffers: env.step_host(ndarray) -> pinned H2D of the actions, kernel, D2H of the
            observations and rewards, every step.
  roofline  HBM: algorithmic bytes per launch / average launch duration of the step kernel over the timed region.
  cpu_baseline  the NumPy oracle (a port of the reference algorithm, oracle/citylearn_oracle.py) on a bounded sample.

`--impl reference` times the CPU implementation of the same path (the oracle port, all host cores via processes; the
reference itself is Python and cannot travel to the GPU box) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / 'oracle'):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

DATASET = 'citylearn_challenge_2022_phase_all'
ENVS_PER_GPU = 4096
METRIC = 'building_env_steps_per_sec'
UNIT = 'building-env steps/s'


def bytes_per_unit(precision: str, n_obs: int, n_act: int, n_buildings: int) -> float:
    """Algorithmic HBM bytes per (building, env) per step: actions read, state read+write, obs + reward +
    district written.  fp64 flow keeps degraded capacity and efficiency as doubles (20 B of state instead of 12 B)."""
    state = 20 if precision == 'fp64' else 12
    return 4 * n_act + 2 * state + 4 * n_obs + 4 + 12.0 / n_buildings


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed regions through NVML in a background thread
    (same fields as the nvidia-smi line in B200_PROFILING.md, without spawning a process that perturbs the host loop)."""

    def __init__(self, index: int, period_s: float = 0.1):
        self.index, self.period = index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.ok = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML enumerates physical GPUs; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = self.index
            if vis:
                try:
                    phys = int(vis.split(',')[self.index])
                except Exception:
                    phys = self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False
            return self._start_smi()
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _start_smi(self):
        """Fallback: the nvidia-smi line of B200_PROFILING.md at a low rate."""
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '250'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            return

        def read():
            for line in self.proc.stdout:
                r = [x.strip() for x in line.split(',')]
                try:
                    self.sm.append(float(r[0]))
                    self.max_mhz = float(r[1])
                    for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[2:6]):
                        if v.lower().startswith('active'):
                            self.reasons.add(name)
                except Exception:
                    pass
        self.thread = threading.Thread(target=read, daemon=True)
        self.thread.start()
        self.ok = True
        self.smi = True

    def _run(self):
        nv = self.nv
        names = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown, 'hw_thermal_slowdown': nv.nvmlClocksEventReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksEventReasonSwThermalSlowdown, 'sw_power_cap': nv.nvmlClocksEventReasonSwPowerCap} \
            if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown') else \
                {'hw_slowdown': nv.nvmlClocksThrottleReasonHwSlowdown, 'hw_thermal_slowdown': nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksThrottleReasonSwThermalSlowdown, 'sw_power_cap': nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def stop(self):
        if not self.ok:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        self._stop.set()
        if getattr(self, 'smi', False):
            self.proc.terminate()
        self.thread.join(timeout=2)
        sm = sorted(self.sm)
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons), 'samples': len(sm)}


def cpu_oracle_rate(n_envs: int, steps: int, seed: int = 0):
    """building-env steps / s of the NumPy oracle on ONE core for `steps` steps of `n_envs` envs."""
    import numpy as np
    from citylearn_b200 import schema as S
    from citylearn_oracle import OracleEnv
    spec = S.load(DATASET)
    env = OracleEnv(spec, n_envs)
    env.reset()
    rng = np.random.RandomState(seed)
    acts = rng.uniform(-1, 1, size=(steps + 1, n_envs, spec.action_dim)).astype('float32')
    env.step(acts[0])
    t0 = time.perf_counter()
    for k in range(steps):
        env.step(acts[k + 1])
    dt = time.perf_counter() - t0
    return spec.n_buildings * n_envs * steps / dt, dt


def effective_cpus() -> int:
    """Host cores this process may really use: min(affinity, cgroup CPU quota)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us'):
        try:
            txt = open(path).read().split()
            if path.endswith('cpu.max'):
                if txt[0] != 'max':
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]))))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open('/sys/fs/cgroup/cpu/cpu.cfs_period_us').read())
                    n = min(n, max(1, q // per))
            break
        except Exception:
            continue
    return n


def _ref_worker(args):
    n_envs, steps, seed = args
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    return cpu_oracle_rate(n_envs, steps, seed)


def run_reference(args):
    """CPU arm: the oracle port on all host cores (one process per core, envs split evenly)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import multiprocessing as mp
    procs = max(1, min(effective_cpus(), 128))
    per = max(1, ENVS_PER_GPU // procs)
    steps, warm = args.steps, args.warmup
    ctx = mp.get_context('fork')
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_ref_worker, [(per, steps + warm, i) for i in range(procs)])
    wall = time.perf_counter() - t0
    # per-process rates exclude construction; sum over processes = whole-host throughput
    value = float(sum(r for r, _ in res))
    ms = 1e3 * max(dt for _, dt in res) / (steps + warm)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps, 'warmup': warm,
        'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{DATASET}: 17 buildings x {per * procs} envs (bounded sample: {procs} processes x {per} envs)',
                   'note': 'CPU port of the reference algorithm (oracle/citylearn_oracle.py, NumPy, float64 intermediates); the Python '
                           'reference itself cannot travel to the GPU box - BASELINE.md has its measured 871.7 building-steps/s/core'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': procs, 'kind': 'port',
                         'sample': f'{procs} processes x {per} envs x {steps + warm} steps, wall {wall:.1f}s'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=200)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'fp32'])
    ap.add_argument('--envs', type=int, default=ENVS_PER_GPU, help='parallel envs per GPU')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    from citylearn_b200 import CityLearnEnv

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    assert torch.cuda.is_available(), 'bench.py needs a CUDA device'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    K, W, E = args.steps, max(args.warmup, 3), args.envs
    env = CityLearnEnv(DATASET, num_envs=E, device=dev, precision=args.precision)
    B, A, L = env.spec.n_buildings, env.spec.action_dim, env._obs_dim
    assert W + K <= env.time_steps - 1, 'steps + warmup must fit in one episode'

    # ---------------- device-resident throughput ----------------
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = torch.rand((W + K, E, A), device=dev, generator=g) * 2 - 1
    obs = torch.empty((K, E, L), device=dev)
    rew = torch.empty((K, E, B), device=dev)
    dst = torch.empty((K, E, 3), device=dev)
    env.reset()
    env.rollout(acts[:W].contiguous(), obs[:W], rew[:W], dst[:W])                 # warm-up steps (untimed)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = env.gpu_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    ev0.record()
    env.rollout(acts[W:], obs, rew, dst)       # EXACTLY K steps
    ev1.record()
    torch.cuda.synchronize(dev)
    ms_total = ev0.elapsed_time(ev1)
    launches = env.gpu_launches - launches0
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    checksum = float(rew.sum().item())            # reads the result back: the step's rewards
    units_per_step = B * E
    value = world * units_per_step * K / (ms_max * 1e-3)

    # ---------------- end to end through the public API with host buffers ----------------
    env.reset()
    host_acts = np.random.RandomState(7 + rank).uniform(-1, 1, size=(W + K, E, A)).astype('float32')
    for k in range(W):
        env.step_host(host_acts[k])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    ev0.record()
    e2e_sum = 0.0
    for k in range(W, W + K):
        o, r, term = env.step_host(host_acts[k])
        e2e_sum += float(r[0, 0])
    ev1.record()
    torch.cuda.synchronize(dev)
    t2 = torch.tensor([ev0.elapsed_time(ev1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_step * K / (float(t2.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    bpu = bytes_per_unit(args.precision, L // B, A // B, B)
    # the timed region is `launches` launches of the rollout kernel (normally one) covering K steps: algorithmic bytes per
    # launch = K/launches steps x bytes per step; duration = CUDA-event time of the region / launches
    steps_per_launch = K / max(launches, 1)
    avg_launch_s = ms_total * 1e-3 / max(launches, 1)
    bytes_per_launch = bpu * units_per_step * steps_per_launch
    achieved = bytes_per_launch / avg_launch_s / 1e9
    traffic = None
    try:
        # measured DRAM bytes (ncu --set full: dram__bytes_read.sum + dram__bytes_write.sum) per step of the same workload, scaled to
        # the steps of one launch like `achieved`; only valid for the default workload the capture was taken on
        rec = json.loads((ROOT / 'profiles' / 'step_kernel_traffic.json').read_text()).get(args.precision)
        if rec and args.envs == ENVS_PER_GPU:
            traffic = (rec['dram_read_bytes_per_step'] + rec['dram_write_bytes_per_step']) * steps_per_launch
    except Exception:
        pass
    cpu = None
    if not args.no_cpu_baseline:
        rate, dt = cpu_oracle_rate(512, 1200)      # ~10-20 s of single-core NumPy work
        cpu = {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
               'sample': f'NumPy oracle, 17 buildings x 512 envs x 1200 steps in {dt:.1f}s on 1 core '
                         f'(reference itself: 871.7 building-steps/s/core, BASELINE.md)'}
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_max / K,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'fp64' else 'f32', 'data': 'synthetic',
        'config': {'workload': f'{DATASET}: {B} buildings x {E} envs per GPU, observations {L}/env, actions {A}/env',
                   'precision': args.precision + (' (float64 intermediates, float32 storage: the reference\'s own flow)' if args.precision == 'fp64' else ''),
                   'mode': 'cl_rollout: ONE persistent kernel launch advances all K steps (state in registers, TMA row ring), actions pre-resident in HBM',
                   'l2': f'every step writes its own obs/reward slab: {K} x {bpu * units_per_step / 1e6:.1f} MB > 126 MB L2',
                   'envs_sharded_across_gpus': True, 'collectives_on_step_path': 0},
        'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': traffic,
                     'kernel': 'advance_kernel', 'bytes_per_unit': bpu, 'bytes_per_step': bpu * units_per_step, 'bytes_per_launch': bytes_per_launch, 'steps_per_launch': steps_per_launch,
                     'avg_launch_us': avg_launch_s * 1e6, 'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback 6650'},
        'cpu_baseline': cpu,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': E * A * 4, 'd2h_bytes_per_step': E * L * 4 + E * B * 4,
                'ms_per_step': float(t2.item()) / K},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'checksum': checksum + e2e_sum * 0.0,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
