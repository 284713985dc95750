#!/usr/bin/env python
"""Segments an ncu source-page CSV of assemble_tiles_kernel at its barriers / role switches and prints, per segment, warp
instructions, stall-sample share and shared-memory wavefronts per tile.  Usage: ncu_segments.py src.csv n_tiles"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1]))); NT = float(sys.argv[2])
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
marks = [i for i, r in enumerate(data) if any(t in r[ix['Source']] for t in ('BAR.', 'USETMAXREG', 'EXIT'))]
tot = sum(num(r, '# Samples') for r in data)
print('warp-instr/tile %.0f  smem wavefronts/tile %.0f (ideal %.0f)' % (sum(num(r, 'Instructions Executed') for r in data) / NT,
      sum(num(r, 'L1 Wavefronts Shared') for r in data) / NT, sum(num(r, 'L1 Wavefronts Shared Ideal') for r in data) / NT))
prev = 0
for m in marks + [len(data)]:
    seg = data[prev:m]
    if seg:
        inst = sum(num(r, 'Instructions Executed') for r in seg); samp = sum(num(r, '# Samples') for r in seg)
        wave = sum(num(r, 'L1 Wavefronts Shared') for r in seg); ideal = sum(num(r, 'L1 Wavefronts Shared Ideal') for r in seg)
        if inst / NT > 5 or samp / tot > 0.005:
            st = defaultdict(float)
            for r in seg:
                for h in stall_cols: st[h] += num(r, h)
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            nxt = data[m][ix['Source']][:34] if m < len(data) else ''
            print(f'[{prev:5d},{m:5d}) inst/tile {inst/NT:6.0f} samples {samp/tot:6.1%} wave/tile {wave/NT:5.0f} ideal {ideal/NT:5.0f} -> {nxt:34s} {[(k[6:], int(v)) for k, v in top]}')
    prev = m
