#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into text: key metrics, stall mix, hottest SASS lines.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, row = r[0], r[1], r[2]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__average_warp_latency_per_inst_issued.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max',
        'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum']
print('kernel:', row[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?')
for i, h in enumerate(hdr):
    if h in keep: print(f'{h:72s} {row[i]:>18s} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]; data = rows[2:]; ix = {h: i for i, h in enumerate(h2)}
def g(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
stalls = [h for h in h2 if h.startswith('stall_') and 'Not' not in h]
agg = {h: sum(g(r, h) for r in data) for h in stalls}; tot = sum(agg.values()) or 1
print('\nSASS lines', len(data), ' warp-instructions executed', int(sum(g(r, 'Instructions Executed') for r in data)), ' samples', int(sum(g(r, '# Samples') for r in data)))
print('warp stall sampling (all samples):')
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:9]: print(f'  {h:26s} {100*v/tot:5.1f} %')
print(f'\nhottest {topn} SASS lines by samples (index, SASS, samples, executed, top stall):')
top = sorted(range(len(data)), key=lambda i: -g(data[i], '# Samples'))[:topn]
for i in sorted(top):
    r = data[i]; st = sorted(((h, g(r, h)) for h in stalls), key=lambda x: -x[1])[0]
    print(f'{i:6d} {r[ix["Source"]][:64]:64s} {int(g(r,"# Samples")):7d} {int(g(r,"Instructions Executed")):9d} {st[0]}')
