#!/usr/bin/env python
"""Raw pinned-memory copy bandwidth of the box (what bounds bench.py's e2e number)."""
import time
import torch
n = 210 * 1024 * 1024 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, src, dst in (("h2d", h, d), ("d2h", d, h)):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms for {n * 4 / 1e6:.0f} MB)")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n // 5, dtype=torch.float32).pin_memory()
d2 = torch.empty(n // 5, dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"duplex: d2h {n * 4 / 1e6:.0f} MB + h2d {n * 4 / 5e6:.0f} MB in {dt * 1e3:.2f} ms")
