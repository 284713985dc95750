"""Developer tool: phase timers of eolc_forces_plan_create (EOLC_PLAN_TIMING=1) on a regular2 n x n sheet.  usage: python scripts/plan_timing.py [n] [shuffle]"""
import os, sys, time
os.environ["EOLC_PLAN_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eol_cloth_b200 as E
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
X, fn = E.meshgen.regular2(n)
N = X.shape[0]
es = E.meshgen.edge_stencils(N, fn)
if len(sys.argv) > 2:
    p = np.random.default_rng(0).permutation(N).astype(np.int32)
    fn = p[fn]; es = np.where(es >= 0, p[np.maximum(es, 0)], -1).astype(np.int32)
    Xn = np.empty_like(X); Xn[p] = X; X = Xn
ctx = E.Context(0)
for r in range(3):
    t = time.perf_counter()
    plan = E.ForcesPlan(ctx, N, fn, es, X_hint=X)
    print("plan_create total %.1f ms" % ((time.perf_counter() - t) * 1e3), file=sys.stderr)
    plan.close()
