#!/usr/bin/env python
"""Developer tool: per-phase clock totals of assemble_tiles_kernel (needs a -DEOLC_TILE_CLOCKS build, scripts/build_variant.sh).
Usage: EOLC_LIB=scratch/variants/libeolc_clocks.so python scripts/tile_clocks.py [n]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import eol_cloth_b200 as E
from eol_cloth_b200 import capi
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ctx = E.Context(0)
X, fn, es, x = bench.make_sheet(n, 0)
plan = E.ForcesPlan(ctx, X.shape[0], fn, es, X_hint=X)
dev = torch.device("cuda", 0)
x_d = torch.from_numpy(x).to(dev); X_d = torch.from_numpy(X.copy()).to(dev)
f_d = torch.empty(3 * X.shape[0], dtype=torch.float64, device=dev)
M_d = torch.empty(plan.nnz[0], dtype=torch.float64, device=dev); K_d = torch.empty(plan.nnz[1], dtype=torch.float64, device=dev)
for _ in range(3):
    plan.fill_dev(x_d.data_ptr(), X_d.data_ptr(), bench.MAT, bench.GRAV, bench.H, f_d.data_ptr(), M_d.data_ptr(), K_d.data_ptr(), n_scenes=1)
torch.cuda.synchronize()
L = capi.lib()
L.eolc_debug_tile_clocks.argtypes = [capi.c_vp, ctypes.c_void_p, ctypes.c_int]
buf = np.zeros((148 * 64, 7), dtype=np.uint64)
rows = L.eolc_debug_tile_clocks(plan.handle, buf.ctypes.data, buf.shape[0])
d = buf[:rows].astype(np.float64)
nw = rows // 148 if rows % 148 == 0 else 8
tiles = d[:, 6]
per = d[:, :6] / tiles[:, None]
names = ["pf|P2", "P1", "bar1", "P3", "bar2", "co|barP23"]
print("rows", rows, "tiles/CTA", tiles.mean(), "clk/tile total", per.sum(1).mean())
print("all warps mean :", " ".join(f"{nm}={v:.0f}" for nm, v in zip(names, per.mean(0))))
for w in range(nw):
    sel = per[w::nw]
    print(f"warp {w:2d}        :", " ".join(f"{nm}={v:.0f}" for nm, v in zip(names, sel.mean(0))))
