#!/usr/bin/env python
"""Per-role cycle counters of alac_decode_kernel on the bench workload (developer tool, needs a GPU)."""
import os, sys, ctypes as C
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from alac_b200_loader import load_package
pkg = load_package()
wl = bench.build_workload(sys.argv[1] if len(sys.argv) > 1 else 'c2', seed=2, threads=bench.host_cores())
n = len(wl['sizes']); dec = pkg.PacketDecoder(pkg.ParseMagicCookie(wl['cookie']), 0)
stride = (dec.frame_bytes + 15) // 16 * 16
d_packed = torch.from_numpy(wl['packed']).cuda(); d_off = torch.from_numpy(wl['offsets'].view(np.int64)).cuda()
d_sz = torch.from_numpy(wl['sizes'].view(np.int32)).cuda(); d_pcm = torch.empty(n * stride, dtype=torch.uint8, device='cuda')
d_nb = torch.zeros(n, dtype=torch.int32, device='cuda'); d_st = torch.zeros(n, dtype=torch.int32, device='cuda')
nct = min((n + 31) // 32, 148 * 8)
buf = torch.zeros(nct * 16, dtype=torch.int64, device='cuda')
import os
if os.environ.get('ALAC_DEBUG_FLAGS'):
    g = pkg.lib.alacb200_debug_flags; g.argtypes=[C.c_uint]; g.restype=C.c_int32; assert g(int(os.environ['ALAC_DEBUG_FLAGS'])) == 0
def run():
    rc = pkg.lib.alacb200_decode_packets_device(dec._h, d_packed.data_ptr(), d_packed.numel(), d_off.data_ptr(), d_sz.data_ptr(), n,
                                                d_pcm.data_ptr(), stride, d_nb.data_ptr(), d_st.data_ptr(), None)
    assert rc == 0; torch.cuda.synchronize()
run()
f = pkg.lib.alacb200_debug_role_cycles; f.argtypes = [C.c_void_p]; f.restype = C.c_int32
assert f(buf.data_ptr()) == 0
run(); run()
f(None)
b = buf.cpu().numpy().reshape(nct, 16).astype(np.float64)
names = ['E total', 'E wait-empty', 'E top-up', 'P total', 'P wait-full', '-', '-', '-', 'tail w0', 'tail w1', '-', 'tag', '-', '-']
for k, nm in enumerate(names):
    if nm in ('tag', '-'): continue
    print(f'{nm:14s} mean {b[:,k].mean()/1e6:8.3f} Mcyc   max {b[:,k].max()/1e6:8.3f} Mcyc')
tag = buf.cpu().numpy().reshape(nct, 16)[:, 11]
smid, wid = tag >> 8, tag & 255
import collections
per_sm = collections.defaultdict(list)
for c in range(nct): per_sm[int(smid[c])].append((int(wid[c]), b[c,0]/1e6, (b[c,0]-b[c,1])/1e6))
raw = buf.cpu().numpy().reshape(nct, 16)
w14 = raw[:, 14].copy().view(np.uint8).reshape(nct, 8)
NW = 2
print('hardware warp slots of CTA warps 0..1 (first 12 CTAs) + entropy SMSP:', [tuple(int(x) for x in w14[c, :5]) for c in range(12)])
print('CTAs whose warps sit on distinct SMSPs:', int(sum(len(set(int(x) % 4 for x in w14[c, :NW])) == NW for c in range(nct))), 'of', nct)
print('SMSP histogram of all role warps:', collections.Counter(int(x) % 4 for c in range(nct) for x in w14[c, :NW]))
print('SMSP histogram of entropy warps:', collections.Counter(int(w14[c, 4]) for c in range(nct)))

cnt = collections.Counter(len(v) for v in per_sm.values()); print('CTAs per SM histogram', dict(cnt))
for k in (1,2,3,4):
    t=[x[1] for v in per_sm.values() if len(v)==k for x in v]
    if t: print(f'  SMs with {k} CTAs: E total mean {np.mean(t):.2f} max {np.max(t):.2f} Mcyc')
slow = sorted(per_sm.items(), key=lambda kv: -max(x[1] for x in kv[1]))[:6]
for sm_, v in slow: print('  slow SM', sm_, [(w, w % 4, round(t,2), round(bz,2)) for w,t,bz in v])
fast = sorted(per_sm.items(), key=lambda kv: max(x[1] for x in kv[1]))[:4]
for sm_, v in fast: print('  fast SM', sm_, [(w, w % 4, round(t,2), round(bz,2)) for w,t,bz in v])
spp = wl['frames'] / n
eb=(b[:,0]-b[:,1])/(2*spp)

print(f'E busy cycles/sample over CTAs: min {eb.min():.0f} p10 {np.percentile(eb,10):.0f} median {np.median(eb):.0f} p90 {np.percentile(eb,90):.0f} p97 {np.percentile(eb,97):.0f} max {eb.max():.0f}')
print(f'per decoded sample (frames/packet = {spp:.0f}): E busy {(b[:,0]-b[:,1]).mean()/(2*spp):.0f} cyc,  P busy {(b[:,3]-b[:,4]).mean()/(2*spp):.0f} cyc')
