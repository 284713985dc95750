#!/usr/bin/env python
"""Developer measurement of the fill's consumer on the device: b = -(M v + h f) and CG iterations on the 1024^2 sheet (HBM-bound
kernels: bytes = values + 4 B per block index + vectors)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eol_cloth_b200 as E
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = E.Context(0)
X, fn, es, x = bench.make_sheet(n, 0)
N = X.shape[0]
plan = E.ForcesPlan(ctx, N, fn, es, X_hint=X)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
xd = torch.from_numpy(x).to(dev); Xd = torch.from_numpy(X.copy()).to(dev)
f = torch.empty(3 * N, dtype=torch.float64, device=dev); M = torch.empty(plan.nnz[0], dtype=torch.float64, device=dev)
K = torch.empty(plan.nnz[1], dtype=torch.float64, device=dev)
v = torch.zeros(3 * N, dtype=torch.float64, device=dev); b = torch.empty_like(v); sol = torch.empty_like(v)
torch.cuda.synchronize()
plan.fill_dev(xd.data_ptr(), Xd.data_ptr(), bench.MAT, bench.GRAV, bench.H, f.data_ptr(), M.data_ptr(), K.data_ptr())
for _ in range(3): plan.rhs_dev(M.data_ptr(), f.data_ptr(), v.data_ptr(), bench.H, b.data_ptr())
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(20): plan.rhs_dev(M.data_ptr(), f.data_ptr(), v.data_ptr(), bench.H, b.data_ptr())
e1.record(stream); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
by = 8 * plan.nnz[0] + 4 * plan.nnz[0] / 9 + 8 * 3 * N * 3
print(f"rhs: {ms:.4f} ms, {by / ms / 1e6:.0f} GB/s algorithmic ({by / 1e6:.0f} MB)")
plan.solve_cg_dev(K.data_ptr(), b.data_ptr(), sol.data_ptr(), tol=1e-30, max_iter=8)
torch.cuda.synchronize()
t = time.perf_counter()
it, res = plan.solve_cg_dev(K.data_ptr(), b.data_ptr(), sol.data_ptr(), tol=1e-30, max_iter=64)
dt = (time.perf_counter() - t) / it * 1e3
by = 8 * plan.nnz[1] + 4 * plan.nnz[1] / 9 + 8 * 3 * N * 12
print(f"cg: {dt:.4f} ms / iteration ({it} iterations, wall clock), {by / dt / 1e6:.0f} GB/s algorithmic ({by / 1e6:.0f} MB / iteration)")
t = time.perf_counter()
it, res = plan.solve_cg_dev(K.data_ptr(), b.data_ptr(), sol.data_ptr(), tol=1e-6, max_iter=4000)
print(f"cg to 1e-6 (cap 4000): {it} iterations, rel. residual {res:.2e}, {(time.perf_counter() - t) * 1e3:.1f} ms")
