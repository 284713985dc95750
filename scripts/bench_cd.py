"""CD narrow phase, device-resident runs, for ncu: (a) CD2 on the 512x512 box scene, (b) the 4096 x 64x64 ensemble batch.
usage: python scripts/bench_cd.py [reps] [scenes]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eol_cloth_b200 as E  # noqa: E402
from eol_cloth_b200.collisions import make_obstacles  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = E.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
THR = E.meshgen.BOX_THRESHOLD


def timed(plan, xd, obs, S, label):
    for _ in range(2):
        off = plan.run_resident(xd.data_ptr(), obs, 0, 0, n_scenes=S)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t = time.perf_counter()
    e0.record(stream)
    for _ in range(reps):
        off = plan.run_resident(xd.data_ptr(), obs, 0, 0, n_scenes=S)
    e1.record(stream)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    pt, ln = plan.stats()
    print(f"{label}: {dt * 1e3:.3f} ms wall, {e0.elapsed_time(e1) / reps:.3f} ms stream, {int(off[-1])} contacts, {ln} launches, {pt / dt / 1e9:.1f} G pair tests/s", flush=True)


X, fn = E.meshgen.regular2(512)
x = E.meshgen.box_scene_state(X, seed=0)
obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame()[None])
plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
xd = torch.from_numpy(x).to(dev)
torch.cuda.synchronize()
timed(plan, xd, obs, 1, "sheet512")
plan.close()

S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
X, fn = E.meshgen.regular2(64)
c = np.array([0.9175, -0.25, -0.549])
obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c)[None])
plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
xs = np.stack([E.meshgen.box_scene_state(X, seed=s, centre=c) for s in range(S)])
xd = torch.from_numpy(xs).to(dev)
torch.cuda.synchronize()
timed(plan, xd, obs, S, "ensemble%dx64" % S)
plan.close()
ctx.close()
