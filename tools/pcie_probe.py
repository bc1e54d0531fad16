import torch, time
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device='cuda')
h2 = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
d2 = torch.empty(2 << 30, dtype=torch.uint8, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name in ('d2h', 'h2d', 'both'):
    for _ in range(3):
        torch.cuda.synchronize(); t = time.perf_counter()
        if name in ('d2h', 'both'):
            with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        if name in ('h2d', 'both'):
            with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(name, 'd2h GB/s', n / dt / 1e9 if name != 'h2d' else 0, 'h2d GB/s', (2 << 30) / dt / 1e9 if name != 'd2h' else 0, 'ms', dt * 1e3)
