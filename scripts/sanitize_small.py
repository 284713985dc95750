"""Small fills (Lagrangian + EOL + batched) for compute-sanitizer: python scripts/sanitize_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import eol_cloth_b200 as E
ctx = E.Context(0)
X, fn = E.meshgen.regular2(40)
es = E.meshgen.edge_stencils(X.shape[0], fn)
x = E.meshgen.drape_state(X, seed=1)
for eol in (None, "line"):
    mesh = dict(x=x, X=X, face_nodes=fn, edge_stencil=es)
    if eol:
        e = np.full(X.shape[0], -1, np.int32); e[np.arange(1, 39) * 40 + 20] = np.arange(38); mesh["eol_index"] = e
    F = E.Forces(ctx)
    for _ in range(2):
        F.fill(mesh, E.Material.DEFAULT, (0.0, 0.0, -9.8), 0.5e-2)
    print("fill ok", eol, F.f.size, float(np.abs(F.MDK[2]).sum()))
plan = E.ForcesPlan(ctx, X.shape[0], fn, es)
print("normals", [a.shape for a in plan.normals(x)])
ctx.close()
