#!/usr/bin/env python
"""Developer statistic: modelled shared-memory wavefronts of the phase-2 block pulls of a plan (tests/hostmath model)."""
import sys, ctypes, numpy as np, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(R, 'tests')); sys.path.insert(0, R)
import eol_cloth_b200 as E
from test_tiles_host import HostTiles
L = ctypes.CDLL(os.path.join(R, 'tests/hostmath/libhostmath.so'))
L.hm_plan_pull_conflicts.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
for gen, n in (('regular2', 256), ('build4', 64)):
    X, fn = getattr(E.meshgen, gen)(n); es = E.meshgen.edge_stencils(X.shape[0], fn)
    T = HostTiles(L, X.shape[0], fn, es, X, True)
    out = (ctypes.c_double * 2)()
    L.hm_plan_pull_conflicts(T.h, out)
    print(gen, n, 'tiles', T.info['n_tiles'], 'pull wavefronts / ideal = %.3f' % (out[0] / out[1]), 'per tile %.1f (ideal %.1f)' % (out[0] / T.info['n_tiles'], out[1] / T.info['n_tiles']))
