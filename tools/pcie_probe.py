#!/usr/bin/env python
"""Measured PCIe floor for the end-to-end leg: pinned H2D of the compressed bytes and D2H of the PCM, alone and together."""
import time, torch
MB = 1 << 20
h_in = torch.empty(235 * MB, dtype=torch.uint8).pin_memory(); d_in = torch.empty_like(h_in, device='cuda')
d_out = torch.empty(346 * MB, dtype=torch.uint8, device='cuda'); h_out = torch.empty(346 * MB, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, n=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
print(f'H2D 235 MiB {a:.2f} ms ({235*MB/a/1e6:.1f} GB/s)  D2H 346 MiB {b:.2f} ms ({346*MB/b/1e6:.1f} GB/s)  both {c:.2f} ms')
