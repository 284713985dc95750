"""Parity of the CUDA narrow phase (through the C ABI): identical count, order and every field BIT FOR BIT against
  * the reference's own code (oracle/_ref/libbtc_ref.so: boxTriCollision.cpp + Collisions.cpp + raytri.cpp compiled unmodified
    against oracle/mini_eigen; the prebuilt library travels to the GPU box) wherever the reference's int edge hash is defined
    (N < ~19 k nodes: all small cases and the 64x64 ensemble scenes),
  * the CPU oracle (oracle/cd_ref.cpp, pinned to that library by tests/test_cd_ref_pin.py) at every size,
  * the golden fixtures (written from libbtc_ref.so)."""
import os

import numpy as np
import pytest

import eol_cloth_b200 as E
from eol_cloth_b200.collisions import make_obstacles
from util import assert_contacts_equal, bit_identical_fraction

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
THR = E.meshgen.BOX_THRESHOLD


def _rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _both(ctx, oracle, fn, x, obs, what):
    mesh = dict(x=x, face_nodes=fn)
    for fnc, flag, remap in ((E.CD, 1, 1), (E.CD2, 0, 0)):
        cls = []
        fnc(ctx, mesh, obs, cls)
        got = np.array(cls, dtype=E.CONTACT_DTYPE) if cls else np.zeros(0, E.CONTACT_DTYPE)
        ref = oracle.cd(fn, x, obs.cdthreshold, obs.pxyz, obs.pnorms, obs.box_whd, obs.box_E, flag, remap)
        assert_contacts_equal(got, ref, what=f"{what} {fnc.__name__}")
        assert got.tobytes() == ref.tobytes(), f"{what} {fnc.__name__}: not bit-identical to the oracle"
        if oracle.ref_hash_defined(len(x), len(fn)):
            own = oracle.ref_cd(fn, x, obs.cdthreshold, obs.pxyz, obs.pnorms, obs.box_whd, obs.box_E, flag)
            assert got.tobytes() == own.tobytes(), f"{what} {fnc.__name__}: not bit-identical to the reference's own code"
    return got, ref


CASES = [
    ("regular2", 24, tuple(E.meshgen.BOX_CENTRE), None, 0),
    ("build4", 16, (0.9175, -0.25, -0.549), None, 1),
    ("regular2", 40, (0.9175, -0.25, -0.549), None, 2),
    ("regular2", 33, (0.7, 0.3, -0.549), ((0, 0, 1), 0.3), 3),          # rotated box, corners + edges under the cloth
    ("build4", 21, (0.5, 0.5, -0.549), ((1, 2, 0.5), 0.05), 4),
]


@pytest.mark.parametrize("gen,n,centre,rot,seed", CASES)
def test_box_scene_matches_oracle(ctx, oracle, gen, n, centre, rot, seed):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
    R = None if rot is None else _rot(*rot)
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(np.asarray(centre), R)[None])
    got, ref = _both(ctx, oracle, fn, x, obs, f"{gen}{n}")
    assert len(ref) > 0


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_pathological_faces_and_borderline_vertices(ctx, oracle, seed):
    """The conservative screens in front of the reference's expressions (plane distance without the normalisation, the division-free
    barycentric verdict, the pair culls) must not change a decision where the geometry is bad: sliver faces (two vertices 1e-9 /
    1e-13 apart, three vertices on a line), vertices and faces exactly in the planes of the box's top and sides, vertices at the
    box's corners and on its edges, faces a hair inside / outside the 5 thr band.  Bit for bit against the reference's own code."""
    rng = np.random.default_rng(100 + seed)
    n = 28
    X, fn = E.meshgen.regular2(n)
    centre = np.array([0.9175, -0.25, -0.549])
    x = E.meshgen.box_scene_state(X, seed=seed, centre=centre)
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(centre)[None])
    lo, hi = centre - 0.5 * E.meshgen.BOX_WHD, centre + 0.5 * E.meshgen.BOX_WHD
    top = hi[2]
    over = np.where((x[:, 0] > lo[0]) & (x[:, 0] < hi[0]) & (x[:, 1] > lo[1]) & (x[:, 1] < hi[1]))[0]
    assert len(over) > 30
    pick = rng.permutation(over)
    # slivers: a vertex almost on its right-hand neighbour, and three vertices of a row on one line
    for k, eps in ((0, 1e-9), (1, 1e-13), (2, 0.0)):
        a = pick[k]
        if (a + 1) % n:
            x[a + 1] = x[a] + eps
    a = pick[3]
    if a + n + 1 < len(x):
        x[a + n + 1] = 2.0 * x[a + n] - x[a + n - 1] if (a + n) % n else x[a + n + 1]
    # exactly in the top plane, a hair above / below it, and at the edge of the 5 thr band
    for k, dz in zip(range(4, 12), (0.0, 1e-16, -1e-16, 5.0 * THR, 5.0 * THR * (1 + 1e-15), 5.0 * THR * (1 - 1e-15), -5.0 * THR, THR)):
        x[pick[k], 2] = top + dz
    # at the box's corners and on its edges (top face), and exactly in a side plane
    cs = [np.array([sx, sy, top]) for sx in (lo[0], hi[0]) for sy in (lo[1], hi[1])]
    near = [int(np.argmin(np.linalg.norm(x[:, :2] - c[:2], axis=1))) for c in cs]
    for a, c in zip(near, cs):
        x[a] = c
    x[pick[12], 0] = hi[0]; x[pick[13], 1] = lo[1]
    x[pick[14]] = np.array([hi[0], x[pick[14], 1], top])
    got, ref = _both(ctx, oracle, fn, x, obs, f"pathological {seed}")
    assert len(ref) > 50


def test_points_and_two_boxes(ctx, oracle):
    X, fn = E.meshgen.regular2(30)
    x = E.meshgen.box_scene_state(X, seed=5, centre=np.array([0.9175, -0.25, -0.549]))
    pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3, x[100] - 2e-3])
    pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0], [0, 0.6, 0.8]])
    whd = np.stack([E.meshgen.BOX_WHD, [0.3, 0.3, 0.3]])
    Em = np.stack([E.meshgen.box_frame(np.array([0.9175, -0.25, -0.549])), E.meshgen.box_frame(np.array([0.15, 0.8, -0.36]))])
    obs = make_obstacles(THR, pxyz, pn, whd, Em)
    got, ref = _both(ctx, oracle, fn, x, obs, "points+2boxes")
    assert {(1, 3), (2, 2), (3, 1)} <= set(zip(ref["count1"].tolist(), ref["count2"].tolist()))


def test_no_contacts_and_no_obstacles(ctx, oracle):
    X, fn = E.meshgen.regular2(8)
    x = np.c_[X, np.full(len(X), 5.0)]
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame()[None])
    cls = []
    E.CD(ctx, dict(x=x, face_nodes=fn), obs, cls)
    assert cls == []
    E.CD2(ctx, dict(x=x, face_nodes=fn), make_obstacles(THR), cls)
    assert cls == []


@pytest.mark.parametrize("name,gen,n,centre,seed,points", [
    ("cd_regular2_n24", "regular2", 24, tuple(E.meshgen.BOX_CENTRE), 0, False),
    ("cd_build4_n16_corner", "build4", 16, (0.9175, -0.25, -0.549), 1, True)])
def test_matches_golden(ctx, name, gen, n, centre, seed, points):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
    pxyz = pn = None
    if points:
        pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3])
        pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]])
    obs = make_obstacles(THR, pxyz, pn, E.meshgen.BOX_WHD[None], E.meshgen.box_frame(np.asarray(centre))[None])
    for key, fnc in (("cd", E.CD), ("cd2", E.CD2)):
        cls = []
        fnc(ctx, dict(x=x, face_nodes=fn), obs, cls)
        got = np.array(cls, dtype=E.CONTACT_DTYPE)
        assert_contacts_equal(got, g[key], what=f"{name} {key}")
        assert bit_identical_fraction(got, g[key]) == 1.0


def test_edge_table_matches_oracle(ctx, oracle):
    X, fn = E.meshgen.build4(11)
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    tab, _ = oracle.cd_edges(fn, np.c_[X, np.zeros(len(X))])
    assert np.array_equal(plan.edge_table(), tab)


def test_box_scene_512_bit_exact(ctx, oracle):
    """BASELINE config 3: simulationSettingsBox.json geometry scaled to a 512x512 cloth, full list comparison."""
    X, fn = E.meshgen.regular2(512)
    x = E.meshgen.box_scene_state(X, seed=0)
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame()[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    got = plan.run(x, obs, 0, 0)
    ref = oracle.cd(fn, x, THR, None, None, obs.box_whd, obs.box_E, 0, 0)
    assert_contacts_equal(got, ref, what="512 box scene")
    nA = int(((ref["count1"] == 3) & (ref["count2"] == 1)).sum())
    assert 0.6 * len(X) < nA < 0.75 * len(X)           # ~0.68 N type-(A) records
    assert int((ref["count1"] == 2).sum()) > 0          # band of type-(C) records along x = 0.3175
    assert got.tobytes() == ref.tobytes()


def test_batched_scenes_equal_single(ctx):
    import torch
    X, fn = E.meshgen.regular2(20)
    c = np.array([0.9175, -0.25, -0.549])
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c)[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    xs = np.stack([E.meshgen.box_scene_state(X, seed=s, centre=c) for s in range(5)])
    xd = torch.from_numpy(xs).to(torch.device("cuda", ctx.device))
    torch.cuda.synchronize()
    allc, off = plan.run(xd.data_ptr(), obs, 0, 0, x_is_device_ptr=True, n_scenes=5)
    for s in range(5):
        single = plan.run(xs[s], obs, 0, 0)
        assert allc[off[s]:off[s + 1]].tobytes() == single.tobytes()


def test_batch_of_20000_small_scenes(ctx):
    """One run takes n_scenes * max(n_boxes, n_points, 1) <= 65535: 20,000 scenes of a 12x12 sheet over the box in one call, sampled
    scenes equal to single-scene runs bit for bit; one scene more than the limit is refused with EOLC_ERR_ARG."""
    import torch
    X, fn = E.meshgen.regular2(12)
    c = np.array([0.9175, -0.25, -0.549])
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c)[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    S = 20000
    base = np.stack([E.meshgen.box_scene_state(X, seed=s, centre=c) for s in range(50)])
    xs = base[np.arange(S) % 50].copy()
    xs[:, :, 2] += 1e-6 * (np.arange(S) // 50)[:, None]          # 20,000 different states
    xd = torch.from_numpy(xs).to(torch.device("cuda", ctx.device))
    torch.cuda.synchronize()
    allc, off = plan.run(xd.data_ptr(), obs, 0, 0, capacity=1_000_000, x_is_device_ptr=True, n_scenes=S)
    assert off[0] == 0 and off[-1] == len(allc) and len(allc) > S
    for s in (0, 1, 49, 50, 8191, 8192, 12345, S - 1):
        single = plan.run(xs[s], obs, 0, 0)
        assert allc[off[s]:off[s + 1]].tobytes() == single.tobytes(), f"scene {s}"
    with pytest.raises(Exception):
        plan.run_resident(xd.data_ptr(), obs, 0, 0, n_scenes=65536)
    plan.close()


def _records_on_device(plan):
    """Copies the records of the plan's last device-resident run to the host."""
    import torch
    from cuda.bindings import runtime as cudart
    torch.cuda.synchronize()                 # pass 2 may still be running on the library's (non-blocking) stream
    ptr, n = plan.contacts_dev()
    out = np.zeros(n, dtype=E.CONTACT_DTYPE)
    if n:
        (err,) = cudart.cudaMemcpy(out.ctypes.data, ptr, out.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
        assert int(err) == 0
    return out


def test_resident_run_speculates_pass_2(ctx, monkeypatch):
    """eolc_cd_run_batched_resident_dev launches pass 2 before the host has seen the totals whenever the plan's record buffers exist
    from an earlier run (every write guarded by their capacities) and repeats it if they were too small.  Same records as the
    host-list run in every case: first run (no buffers yet), a run that fits, a run with ~30x the contacts (buffers too small), a
    run with none, and with the speculation switched off."""
    import torch
    X, fn = E.meshgen.regular2(32)
    c = np.array([0.9175, -0.25, -0.549])
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c)[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    dev = torch.device("cuda", ctx.device)
    S = 6
    draped = np.stack([E.meshgen.box_scene_state(X, seed=s, centre=c) for s in range(S)])
    few = draped.copy(); few[1:, :, 2] += 5.0                       # only scene 0 touches the box
    none = draped.copy(); none[:, :, 2] += 5.0
    order = [("first: few", few), ("fits: few again", few), ("grows: all draped", draped), ("fits: all draped", draped),
             ("none", none), ("few after none", few)]
    for what, xs in order:
        xd = torch.from_numpy(xs).to(dev)
        torch.cuda.synchronize()
        off = plan.run_resident(xd.data_ptr(), obs, 0, 0, n_scenes=S)
        got = _records_on_device(plan)
        monkeypatch.setenv("EOLC_CD_NO_SPECULATION", "1")
        off2 = plan.run_resident(xd.data_ptr(), obs, 0, 0, n_scenes=S)
        got2 = _records_on_device(plan)
        monkeypatch.delenv("EOLC_CD_NO_SPECULATION")
        assert np.array_equal(off, off2) and got.tobytes() == got2.tobytes(), what
        assert off[-1] == len(got), what
        for s_ in (0, S - 1):
            single = E.CollisionPlan(ctx, X.shape[0], fn, THR)
            one = single.run(xs[s_], obs, 0, 0)
            single.close()
            assert got[off[s_]:off[s_ + 1]].tobytes() == one.tobytes(), f"{what}: scene {s_}"
    assert len(got) > 0
    plan.close()


def test_resident_speculation_with_points_and_two_boxes(ctx):
    """The same with every section present (obstacle points with the EOL flag, two boxes, CD's index remap): second and third runs
    on the plan speculate; records equal to the host-list run."""
    import torch
    X, fn = E.meshgen.regular2(30)
    dev = torch.device("cuda", ctx.device)
    states = [E.meshgen.box_scene_state(X, seed=s, centre=np.array([0.9175, -0.25, -0.549])) for s in (5, 6, 7)]
    x0 = states[0]
    pxyz = np.array([[0.25, 0.25, x0[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x0[5] + 1e-3, x0[100] - 2e-3])
    pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0], [0, 0.6, 0.8]])
    whd = np.stack([E.meshgen.BOX_WHD, [0.3, 0.3, 0.3]])
    Em = np.stack([E.meshgen.box_frame(np.array([0.9175, -0.25, -0.549])), E.meshgen.box_frame(np.array([0.15, 0.8, -0.36]))])
    obs = make_obstacles(THR, pxyz, pn, whd, Em)
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    ref_plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    for flag, remap in ((1, 1), (0, 0)):
        for x in states:
            xd = torch.from_numpy(x).to(dev)
            torch.cuda.synchronize()
            off = plan.run_resident(xd.data_ptr(), obs, flag, remap)
            got = _records_on_device(plan)
            ref = ref_plan.run(x, obs, flag, remap)
            assert off[-1] == len(ref) and got.tobytes() == ref.tobytes()
    assert {(1, 3), (2, 2), (3, 1)} <= set(zip(ref["count1"].tolist(), ref["count2"].tolist()))
    plan.close(); ref_plan.close()


def test_ensemble_4096_scenes_sampled_against_reference(ctx, oracle):
    """BASELINE configs[4]: the full batch of 4096 independent 64x64 scenes (state seed = scene id) against the box, one batched
    call per chunk; 40 sampled scenes are compared bit for bit with the reference's own code (libbtc_ref.so) and the oracle."""
    import torch
    n, S, chunk = 64, 4096, 512
    X, fn = E.meshgen.regular2(n)
    c = np.array([0.9175, -0.25, -0.549])
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c)[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    sample = sorted(set(np.random.default_rng(0).choice(S, 38, replace=False).tolist()) | {0, S - 1})
    total = 0
    for s0 in range(0, S, chunk):
        xs = np.stack([E.meshgen.box_scene_state(X, seed=s, centre=c) for s in range(s0, s0 + chunk)])
        xd = torch.from_numpy(xs).to(torch.device("cuda", ctx.device))
        torch.cuda.synchronize()
        allc, off = plan.run(xd.data_ptr(), obs, 0, 0, x_is_device_ptr=True, n_scenes=chunk)
        total += len(allc)
        assert off[0] == 0 and off[-1] == len(allc) and (np.diff(off) > 0).all()
        for s in sample:
            if s0 <= s < s0 + chunk:
                got = allc[off[s - s0]:off[s - s0 + 1]]
                own = oracle.ref_cd(fn, xs[s - s0], THR, None, None, obs.box_whd, obs.box_E, 0)
                assert got.tobytes() == own.tobytes(), f"scene {s}: not bit-identical to the reference's own code"
                assert got.tobytes() == oracle.cd(fn, xs[s - s0], THR, None, None, obs.box_whd, obs.box_E, 0, 0).tobytes()
    assert total > 1000 * S


def test_pair_list_overflow_repeats_the_pass(ctx, oracle, monkeypatch):
    """Section C's pair work list is sized by a heuristic; when it overflows the device has counted the pairs and the host repeats
    the pass with the exact size.  Forced here with a 3-entry list: same contacts, bit for bit."""
    X, fn = E.meshgen.regular2(40)
    c = np.array([0.9175, -0.25, -0.549])
    x = E.meshgen.box_scene_state(X, seed=2, centre=c)
    obs = make_obstacles(THR, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c)[None])
    ref = oracle.cd(fn, x, THR, None, None, obs.box_whd, obs.box_E, 0, 0)
    assert int((ref["count1"] == 2).sum()) > 3
    monkeypatch.setenv("EOLC_CD_PAIR_CAP", "3")
    plan = E.CollisionPlan(ctx, X.shape[0], fn, THR)
    got = plan.run(x, obs, 0, 0)
    assert got.tobytes() == ref.tobytes()
    monkeypatch.delenv("EOLC_CD_PAIR_CAP")
    assert plan.run(x, obs, 0, 0).tobytes() == ref.tobytes()
