#!/usr/bin/env python
"""List the hot loops of alac_decode_kernel in a built library: size, VIMNMX / local-memory / move counts.
Usage: python tools/dev/sass_loops.py [path/to/libalacb200.so] [--dump START_HEX]"""
import re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith('--') else 'saprobe-alac_b200/libalacb200.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
ins = []
for l in txt.split('\n'):
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print('instructions', len(ins), 'STL', sum('STL' in t for _, t in ins), 'LDL', sum('LDL' in t for _, t in ins))
if '--dump' in sys.argv:
    a0 = int(sys.argv[sys.argv.index('--dump') + 1], 16)
    on = False
    for a, t in ins:
        if a == a0: on = True
        if on:
            print(hex(a), t)
            m2 = re.search(r'BRA.*0x([0-9a-f]+)', t)
            if m2 and int(m2.group(1), 16) == a0: break
    sys.exit(0)
for idx, (a, t) in enumerate(ins):
    m2 = re.search(r'BRA.*0x([0-9a-f]+)', t)
    if m2:
        tgt = int(m2.group(1), 16)
        if tgt < a and (a - tgt) // 16 < 400:
            body = [x for x in ins if tgt <= x[0] <= a]
            nv = sum('VIMNMX' in b[1] for b in body); flo = sum('FLO' in b[1] for b in body)
            if nv >= 20 or flo >= 2:
                print(hex(tgt), hex(a), 'len', len(body), 'VIMNMX', nv, 'FLO', flo, 'local', sum(('STL' in b[1] or 'LDL' in b[1]) for b in body),
                      'MOV', sum('IMAD.MOV' in b[1] or b[1].startswith('MOV') for b in body), 'LDS', sum('LDS' in b[1] for b in body))
