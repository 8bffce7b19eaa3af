#!/usr/bin/env python
"""Where the local-memory (spill) instructions of each decode kernel build sit relative to the predictor's sample loops
and the ring-slot loops around them. Usage: python tools/dev/spill_sites.py [lib.so]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else 'saprobe-alac_b200/libalacb200.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
funcs, cur = {}, None
for l in txt.split('\n'):
    m = re.search(r'Function : (\S+)', l)
    if m:
        cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if 'decode_kernel' not in name: continue
    print('==', 'lat build' if 'lat' in name else 'throughput build', len(ins), 'instructions, STL', sum('STL' in t for _, t in ins), 'LDL', sum('LDL' in t for _, t in ins))
    loops = []
    for a, t in ins:
        m2 = re.search(r'BRA.*0x([0-9a-f]+)', t)
        if m2 and int(m2.group(1), 16) < a: loops.append((int(m2.group(1), 16), a))
    for lo, hi in loops:
        body = [x for x in ins if lo <= x[0] <= hi]
        nv = sum('VIMNMX' in b[1] for b in body); flo = sum('FLO' in b[1] for b in body)
        if len(body) < 400 and (nv >= 20 or flo >= 4):
            enc = sorted([(h - l, l, h) for l, h in loops if l <= lo and hi <= h and (h - l) > (hi - lo)])
            cl = enc[0] if enc else None
            inner = [t for a, t in ins if lo <= a <= hi and ('LDL' in t or 'STL' in t)]
            outer = [t for a, t in ins if cl and cl[1] <= a <= cl[2] and not (lo <= a <= hi) and ('LDL' in t or 'STL' in t)]
            print('  %s loop %#x len %d: local inside %d %s | enclosing loop len %s: local %d' % ('predictor' if nv >= 20 else 'entropy', lo, len(body), len(inner), inner[:2], cl[0] // 16 if cl else None, len(outer)))
