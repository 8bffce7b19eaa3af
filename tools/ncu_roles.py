#!/usr/bin/env python
"""Per-role instruction counts of alac_decode_kernel from an .ncu-rep captured with --import-source on.
Usage: python tools/ncu_roles.py gpurun_out/prof.ncu-rep channel_samples_per_launch
Roles are told apart by the source line of each SASS instruction (function boundaries read from the .cuh); helper
lines (inline asm wrappers at the top of the file) inherit the role of the instruction before them."""
import collections, csv, io, os, re, subprocess, sys
rep = sys.argv[1]
samples = float(sys.argv[2]) if len(sys.argv) > 2 else 0
# the kernel source as it was when the capture was made travels inside the report (--import-source on)
emb = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda'], capture_output=True, text=True).stdout)))
lines, take = {}, False
for r in emb:
    if not r: continue
    if r[0] == 'File Path' or r[0] == 'File Name': take = r[1].endswith('alac_kernels.cuh'); continue
    if take and r[0].isdigit(): lines[int(r[0])] = r[1]
def line_of(*pats):
    for pat in pats:
        for i in sorted(lines):
            if re.search(pat, lines[i]): return i
    raise KeyError(pats)
marks = [(line_of(r'^struct Entropy \{'), 'E.slow'), (line_of(r'^template <bool QUIET>', r'decode_batch\('), 'E.batch'), (line_of(r'void produce_stream\('), 'E.stream'),
         (line_of(r'^struct ElemHdr'), 'E.parse'), (line_of(r'int32_t delta_step'), 'EMIT'),
         (line_of(r'int32_t sext_bits\(', r'^template <int T, bool MODE'), 'P.reg'), (line_of(r'__noinline__ void stream_generic'), 'P.generic'),
         (line_of(r'void stream_escape_pair\(', r'void predictor_warp\('), 'P.warp'), (line_of(r'^struct EmitArgs'), 'TAIL'), (line_of(r'^__global__ void'), 'KERNEL')]
marks.sort()
def role_of_line(l):
    if l is None or l < marks[0][0]: return None
    r = None
    for ln, name in marks:
        if l >= ln: r = name
    return r
cs = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout)))
addr2line, cur, infile = {}, None, True
for r in cs:
    if not r: continue
    if r[0] == 'File Path': infile = r[1].endswith('alac_kernels.cuh'); continue
    if r[0] in ('Function Name', 'Line No', 'Kernel Name'): continue
    if r[0] != '': cur = int(r[0]) if infile and r[0].isdigit() else None; continue
    if len(r) > 2 and r[2].startswith('0x'): addr2line[int(r[2], 16)] = cur
s = list(csv.reader(io.StringIO(subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout)))
h2 = s[1]; ix = {h: i for i, h in enumerate(h2)}
ALU = ('SEL', 'LOP3', 'SHF', 'IADD3', 'VIADD', 'VIMNMX', 'ISETP', 'PRMT', 'LEA', 'PLOP3', 'SGXT', 'IABS', 'MOV', 'BMSK', 'POPC', 'VABSDIFF')
FMA = ('IMAD', 'FFMA', 'FMUL')
instr, samp, alu, fma, spin = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
last = 'KERNEL'
for r in s[2:]:
    a = int(r[0], 16); sass = r[1].strip(); ex = float(r[ix['Instructions Executed']] or 0); sm = float(r[ix['# Samples']] or 0)
    ro = role_of_line(addr2line.get(a))
    if ro: last = ro
    op = sass.split()[1] if sass.startswith('@') else sass.split()[0]
    instr[last] += ex; samp[last] += sm
    if 'TRYWAIT' in sass or 'NANOSLEEP' in sass or (op == 'BRA' and ex > 5e6): spin[last] += ex
    if op.startswith(ALU) and not op.startswith('IMAD'): alu[last] += ex
    if op.startswith(FMA): fma[last] += ex
tot, ts = sum(instr.values()), sum(samp.values()) or 1
steps = samples / 32 if samples else 0
print(f'warp-instructions executed {tot/1e6:.1f} M' + (f'  = {tot/steps:.0f} per 32-lane sample step' if steps else ''))
print(f'{"role":10s} {"instr M":>9s} {"share":>6s} {"per step":>9s} {"ALU-pipe":>9s} {"FMA-pipe":>9s} {"wait polls":>10s} {"stall samples":>13s}')
for k, v in instr.most_common():
    per = (lambda x: f'{x/steps:9.1f}') if steps else (lambda x: f'{x/1e6:8.1f}M')
    print(f'{k:10s} {v/1e6:9.1f} {100*v/tot:5.1f}% {per(v)} {per(alu[k])} {per(fma[k])} {per(spin[k])} {100*samp[k]/ts:12.1f}%')
print(f'{"all":10s} {tot/1e6:9.1f} {"":6s} {per(tot)} {per(sum(alu.values()))} {per(sum(fma.values()))} {per(sum(spin.values()))}')
