#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (profiles/r02/sass_features.txt).  usage: python scripts/sass_features.py > profiles/r02/sass_features.txt"""
import os, re, subprocess
from collections import Counter, OrderedDict
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "eol_cloth_b200", "libeolc_b200.so")
elf = subprocess.run(["cuobjdump", "-lelf", so], capture_output=True, text=True).stdout
archs = sorted(set(re.findall(r"sm_\d+a?", elf)))
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
WATCH = ["USETMAXREG", "LDGSTS", "UBLKCP", "FENCE.VIEW.ASYNC", "BAR.SYNC", "BAR.ARV", "VOTE", "POPC", "SHFL", "ATOMG", "RED", "STS.64", "STS.128", "LDS.64", "LDS.128",
         "DADD", "DMUL", "DFMA", "MUFU.RCP64H", "MUFU.RSQ64H"]
kernels = OrderedDict()
cur = None
for l in txt.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        cur = m.group(1); kernels[cur] = [0, Counter()]; continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if cur and m:
        kernels[cur][0] += 1
        op = m.group(1)
        for w in WATCH:
            if op == w or op.startswith(w + "."):
                kernels[cur][1][w] += 1
print("# SASS evidence (cuobjdump -sass eol_cloth_b200/libeolc_b200.so, built by eol_cloth_b200/csrc/Makefile for sm_100a only; scripts/sass_features.py)\n")
print("cubin architectures in the library:", archs, "\n")
for k, (n, c) in kernels.items():
    short = re.sub(r"^_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{2}", "", k)
    print(f"{short[:62]:62s} {n:5d} instr  " + "  ".join(f"{w}:{c[w]}" for w in WATCH if c[w]))
print("\nUBLKCP = cp.async.bulk (bulk copy engine, shared -> global); USETMAXREG = setmaxnreg; LDGSTS = cp.async; FENCE.VIEW.ASYNC = fence.proxy.async;")
print("BAR.ARV = bar.arrive (named barriers).  No HMMA / UTCMMA: nothing on this path is a dense contraction (SURVEY §8d).")
