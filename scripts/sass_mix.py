#!/usr/bin/env python
"""Static SASS instruction mix of one kernel of libeolc_b200.so, split at its barriers.  Usage: sass_mix.py <kernel-substring>"""
import re, subprocess, sys
from collections import Counter
so = "eol_cloth_b200/libeolc_b200.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur = None; ops = []
for l in txt.splitlines():
    m = re.search(r'Function : (\S+)', l)
    if m: cur = m.group(1); continue
    if cur and sys.argv[1] in cur:
        m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', l)
        if m: ops.append(m.group(3))
bars = [i for i, o in enumerate(ops) if o.startswith('BAR')]
print('total', len(ops), 'barriers at', bars)
for a, b in zip([0] + bars, bars + [len(ops)]):
    c = Counter('IMAD.MOV' if o.startswith('IMAD.MOV') else o.split('.')[0] for o in ops[a:b])
    f = c['DFMA'] + c['DMUL'] + c['DADD']
    print(f'[{a},{b}) n={b-a} fp64={f}', c.most_common(12))
