#!/usr/bin/env python3
"""Raw PCIe numbers of the box, to put the end-to-end (host-array) figure of bench.py in context: pinned H2D of the
256 MB ray batch alone, D2H of the 128 MB results alone, and both at once on two streams."""
import time, torch
n = 8_000_000
h_in = torch.empty(n * 8, dtype=torch.float32).pin_memory(); h_out = torch.empty(n * 4, dtype=torch.float32).pin_memory()
d_in = torch.empty(n * 8, dtype=torch.float32, device='cuda'); d_out = torch.empty(n * 4, dtype=torch.float32, device='cuda')
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3
def up():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def down():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    up(); down()
for _ in range(3): both()
a, b, c = t(up), t(down), t(both)
print(f"H2D 256 MB: {a:.3f} ms = {0.256/a*1e3:.1f} GB/s; D2H 128 MB: {b:.3f} ms = {0.128/b*1e3:.1f} GB/s; both at once: {c:.3f} ms -> ceiling {n/c/1e3:.0f} Mrays/s")
