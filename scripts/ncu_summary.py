#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) the way profiles/*.md quote it.  Usage: ncu_summary.py file.ncu-rep [top_n]"""
import csv, subprocess, sys, io
from collections import Counter
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 18
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, u = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__average_warps_issue_stalled', 'launch__shared_mem_per_block_dynamic', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max', 'smsp__cycles_active.avg',
        'l1tex__t_sector_hit_rate.pct']
for v in rows[2:]:
    print("== kernel:", v[h.index('Kernel Name')][:80])
    for i, n in enumerate(h):
        if any(n == k or (k.startswith('smsp__average_warps_issue_stalled') and n.startswith(k) and n.endswith('per_issue_active.ratio')) for k in keys):
            print(f"  {n:90s} {u[i]:12s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]; ix = {n: i for i, n in enumerate(h)}
tot = sum(int(r[ix['# Samples']]) for r in data)
mix = Counter(); exe = Counter()
for r in data:
    op = [o for o in r[ix['Source']].split() if not o.startswith('@')][0].split('.')[0]
    mix[op] += int(r[ix['# Samples']]); exe[op] += int(r[ix['Instructions Executed']])
te = sum(exe.values())
print(f"== source page: {len(data)} SASS instr, {te} warp-instr executed, {tot} samples")
for op, c in exe.most_common(14): print(f"  {op:8s} {c:12d} {100*c/te:5.1f}% of instr, {100*mix[op]/tot:5.1f}% of samples")
print("== top stall sites")
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:topn]:
    st = {k: int(r[ix[k]]) for k in h if k.startswith('stall_') and 'Not' not in k and r[ix[k]] not in ('', '0')}
    print(f"  {int(r[ix['# Samples']]):7d} {100*int(r[ix['# Samples']])/tot:5.1f}%  {r[ix['Source']][:58]:58s} {sorted(st.items(), key=lambda x: -x[1])[:2]}")
