#!/usr/bin/env python
"""Measured PCIe floor for the end-to-end leg: pinned H2D of the compressed bytes and D2H of the PCM, alone and together.
Alone or under torchrun (one rank per GPU, all ranks copying at the same time): prints per-rank times and the max.
    python tools/pcie_probe.py [h2d_MiB d2h_MiB]        (defaults: c2's 235 / 346; c3 per rank at N ranks: 2672/N, 3955/N)"""
import os, sys, time, torch
MB = 1 << 20
rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
dev = int(os.environ.get('LOCAL_RANK', 0)) % torch.cuda.device_count()
torch.cuda.set_device(dev)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', dev))
a_mb, b_mb = (int(float(sys.argv[1])), int(float(sys.argv[2]))) if len(sys.argv) > 2 else (235, 346)
h_in = torch.empty(a_mb * MB, dtype=torch.uint8).pin_memory(); d_in = torch.empty_like(h_in, device='cuda')
d_out = torch.empty(b_mb * MB, dtype=torch.uint8, device='cuda'); h_out = torch.empty(b_mb * MB, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def sync():
    torch.cuda.synchronize()
    if dist is not None: dist.barrier(); torch.cuda.synchronize()
def t(fn, n=8):
    fn(); sync(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n * 1e3
    if dist is not None:
        x = torch.tensor([dt], device='cuda'); dist.all_reduce(x, op=dist.ReduceOp.MAX); return dt, float(x.item())
    return dt, dt
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = t(h2d), t(d2h), t(both)
if rank == 0:
    print(f'PCIE_FLOOR ranks={world} per-rank H2D {a_mb} MiB: {a[1]:.2f} ms ({a_mb*MB/a[1]/1e6:.1f} GB/s/rank)  D2H {b_mb} MiB: {b[1]:.2f} ms ({b_mb*MB/b[1]/1e6:.1f} GB/s/rank)  both {c[1]:.2f} ms (max over ranks, all ranks copying at once)')
if dist is not None:
    dist.barrier(); dist.destroy_process_group()
