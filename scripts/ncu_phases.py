#!/usr/bin/env python
"""Per-phase summary of an ncu source-page CSV of assemble_tiles_kernel (split at its barriers).
Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python scripts/ncu_phases.py src.csv [n_tiles] [detail_phase_index]"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
ntiles = float(sys.argv[2]) if len(sys.argv) > 2 else 32768.0
hdr = rows[1]; data = rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def num(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
bars = [i for i, r in enumerate(data) if 'BAR.SYNC' in r[ix['Source']]]
tot_inst = sum(num(r, 'Instructions Executed') for r in data); tot_samp = sum(num(r, '# Samples') for r in data)
print('instr', len(data), 'barriers at', bars, 'warp-instr/tile %.0f' % (tot_inst / ntiles), 'smem wavefronts/tile %.0f' % (sum(num(r, 'L1 Wavefronts Shared') for r in data) / ntiles))
bounds = [0] + bars + [len(data)]
for pi, (a, b) in enumerate(zip(bounds[:-1], bounds[1:])):
    seg = data[a:b]
    inst = sum(num(r, 'Instructions Executed') for r in seg); samp = sum(num(r, '# Samples') for r in seg)
    wave = sum(num(r, 'L1 Wavefronts Shared') for r in seg); ideal = sum(num(r, 'L1 Wavefronts Shared Ideal') for r in seg)
    print(f'phase {pi} [{a},{b}) inst/tile {inst/ntiles:.0f} ({inst/tot_inst:.1%}) samples {samp/tot_samp:.1%} wavefronts/tile {wave/ntiles:.0f} ideal {ideal/ntiles:.0f}')
if len(sys.argv) > 3:
    pi = int(sys.argv[3]); a, b = bounds[pi], bounds[pi + 1]
    agg = defaultdict(lambda: [0, 0, 0, 0])
    for r in data[a:b]:
        s = [t for t in r[ix['Source']].split() if not t.startswith('@')]
        op = s[0] if s else '?'
        g = agg[op]; g[0] += num(r, 'Instructions Executed'); g[1] += num(r, 'L1 Wavefronts Shared'); g[2] += num(r, 'L1 Wavefronts Shared Ideal'); g[3] += num(r, '# Samples')
    for op, g in sorted(agg.items(), key=lambda kv: -kv[1][3])[:22]:
        print(f'  {op:28s} inst/tile {g[0]/ntiles:7.1f} wave/tile {g[1]/ntiles:7.1f} ideal {g[2]/ntiles:7.1f} samples {g[3]:.0f}')
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tot = defaultdict(float)
    for r in data[a:b]:
        for h in stall_cols: tot[h] += num(r, h)
    print('  stalls:', ', '.join(f'{h[6:]}={v:.0f}' for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]))
