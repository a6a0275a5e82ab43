#!/usr/bin/env python
"""Developer probe: copy bandwidth as a function of working-set size (L2-resident vs HBM)."""
import torch
for mb in (4, 8, 16, 32, 48, 64, 96, 128, 256, 1024):
    n = mb * (1 << 20) // 8
    a = torch.rand(n, 2, device="cuda")
    b = torch.empty_like(a)
    for _ in range(5): b.copy_(a)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = max(10, 4096 // mb)
    s.record()
    for _ in range(it): b.copy_(a)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / it
    print(f"copy {mb:5d} MiB src (+same dst): {ms*1e3:9.1f} us  {2*mb*1.048576/ms:8.1f} GB/s", flush=True)
