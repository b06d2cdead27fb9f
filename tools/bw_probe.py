import torch, time
for mb in (8, 64):
    n = mb * 1024 * 1024 // 4
    d = torch.empty(n, device='cuda'); h = torch.empty(n).pin_memory()
    for direction in ('d2h', 'h2d'):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20):
            (h.copy_(d, non_blocking=True) if direction == 'd2h' else d.copy_(h, non_blocking=True))
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print(mb, 'MB', direction, f'{mb/1024/dt:.1f} GB/s', f'{dt*1e6:.0f} us')
