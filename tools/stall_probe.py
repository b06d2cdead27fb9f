"""Are the ~70 ms stalls of CUDA API calls periodic (an external poller holding the driver lock)?"""
import subprocess, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
print(subprocess.run("ps -eo pid,etimes,args | grep -i -E 'nvidia|dcgm|smi|nvml|monitor' | grep -v grep | head -20", shell=True, capture_output=True, text=True).stdout)
d = torch.empty(1024, device='cuda'); h = torch.empty(1024).pin_memory()
pc = time.perf_counter
t_start = pc(); stalls = []
n = 0
while pc() - t_start < 3.0:
    t0 = pc(); h.copy_(d, non_blocking=True); torch.cuda.synchronize(); dt = pc() - t0
    if dt > 2e-3: stalls.append((round((t0 - t_start) * 1e3, 1), round(dt * 1e3, 1)))
    n += 1
print('iterations', n, 'stalls (t_ms, dur_ms):', stalls[:40])
# pure host loop: does the CPU itself stall (hypervisor)?
t_start = pc(); hs = []; last = pc()
while pc() - t_start < 1.0:
    now = pc()
    if now - last > 2e-3: hs.append((round((now - t_start) * 1e3, 1), round((now - last) * 1e3, 1)))
    last = now
print('host-only gaps > 2ms:', hs[:20])
print(subprocess.run("ps -eo pid,etimes,args | grep -i -E 'nvidia|dcgm|smi' | grep -v grep | head", shell=True, capture_output=True, text=True).stdout)
