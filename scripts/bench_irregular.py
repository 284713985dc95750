#!/usr/bin/env python
"""Developer measurement: Forces::fill on a mesh with a RANDOM node / face numbering (worst case for the plan: every tile its own
template, one copy-out run per node) against the same mesh in its structured numbering."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eol_cloth_b200 as E
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ctx = E.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
X, fn = E.meshgen.regular2(n)
for label in ("structured", "shuffled"):
    if label == "shuffled":
        rng = np.random.default_rng(5)
        perm = rng.permutation(X.shape[0])
        Xn = np.empty_like(X); Xn[perm] = X
        X, fn = Xn, perm[fn].astype(np.int32)[rng.permutation(len(fn))]
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=1)
    import time
    t = time.perf_counter()
    plan = E.ForcesPlan(ctx, X.shape[0], fn, es, X_hint=X)
    tp = time.perf_counter() - t
    N = X.shape[0]
    xd = torch.from_numpy(x).to(dev); Xd = torch.from_numpy(X.copy()).to(dev)
    f = torch.empty(3 * N, dtype=torch.float64, device=dev); M = torch.empty(plan.nnz[0], dtype=torch.float64, device=dev)
    K = torch.empty(plan.nnz[1], dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for _ in range(3): plan.fill_dev(xd.data_ptr(), Xd.data_ptr(), bench.MAT, bench.GRAV, bench.H, f.data_ptr(), M.data_ptr(), K.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(10): plan.fill_dev(xd.data_ptr(), Xd.data_ptr(), bench.MAT, bench.GRAV, bench.H, f.data_ptr(), M.data_ptr(), K.data_ptr())
    e1.record(stream); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    el = plan.n_faces + plan.n_interior_edges
    print(f"{label:10s} n={n}: plan {tp:.2f} s, fill {ms:.4f} ms, {el / ms / 1e6:.2f} G elements/s, checksum {float(K.sum().item()):.6e}")
