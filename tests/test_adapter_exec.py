"""The reference-side adapter, EXECUTED: adapter/Collisions_b200.cpp — the file a maintainer drops in place of src/Collisions.cpp —
compiled against the reference's own headers (Collisions.h, external/ArcSim/mesh.hpp, Obstacles.h, Box.h, Points.h, boxTriCollision.h) with
oracle/mini_eigen standing in for Eigen, linked with the reference's boxTriCollision.cpp, include/eolc_host.hpp and libeolc_b200.so
(oracle/_ref/libadapter_cd.so, built by oracle/Makefile; travels prebuilt to the GPU box).  The call goes
    ArcSim Mesh + shared_ptr<Obstacles>  ->  CD / CD2 (adapter)  ->  eolc::host::flatten / CD  ->  C ABI  ->  GPU
and comes back as vector<shared_ptr<btc::Collision>>.  It must equal the reference's own CD / CD2 (libbtc_ref.so) in every field."""
import numpy as np
import pytest

import eol_cloth_b200 as E

pytestmark = pytest.mark.gpu
THR = E.meshgen.BOX_THRESHOLD
C3B = np.array([0.9175, -0.25, -0.549])


def _rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _same(a, b, what):
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} contacts"
    for f in a.dtype.names:
        assert a[f].tobytes() == b[f].tobytes(), f"{what}: field {f} differs"


@pytest.mark.parametrize("gen,n,centre,rot,seed", [("regular2", 40, C3B, None, 2), ("build4", 21, (0.5, 0.5, -0.549), ((1, 2, 0.5), 0.05), 4),
                                                   ("regular2", 64, C3B, None, 7), ("regular2", 33, (0.7, 0.3, -0.549), ((0, 0, 1), 0.3), 3)])
def test_adapter_cd_equals_reference(oracle, gen, n, centre, rot, seed):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
    whd = E.meshgen.BOX_WHD[None]
    Em = E.meshgen.box_frame(np.asarray(centre), None if rot is None else _rot(*rot))[None]
    for which in (1, 0):
        ref = oracle.ref_cd(fn, x, THR, None, None, whd, Em, which)
        got = oracle.ref_cd(fn, x, THR, None, None, whd, Em, which, adapter=True)
        _same(ref, got, f"{gen}{n} {'CD' if which else 'CD2'}")
        assert len(ref) > 0


def test_adapter_points_two_boxes_and_repeated_calls(oracle):
    """Points + two boxes (CD's index remap), then the same mesh again with moved positions (the adapter's cached plan and page-locked
    buffers are reused), then another mesh (the plan is rebuilt)."""
    X, fn = E.meshgen.regular2(30)
    x = E.meshgen.box_scene_state(X, seed=5, centre=C3B)
    pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3, x[100] - 2e-3])
    pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0], [0, 0.6, 0.8]])
    whd = np.stack([E.meshgen.BOX_WHD, [0.3, 0.3, 0.3]])
    Em = np.stack([E.meshgen.box_frame(C3B), E.meshgen.box_frame(np.array([0.15, 0.8, -0.36]), _rot((0.3, 1, 0.2), 0.4))])
    for xs in (x, x + [0.0, 0.0, 3e-4]):
        for which in (1, 0):
            _same(oracle.ref_cd(fn, xs, THR, pxyz, pn, whd, Em, which), oracle.ref_cd(fn, xs, THR, pxyz, pn, whd, Em, which, adapter=True), "points+2boxes")
    X2, fn2 = E.meshgen.build4(14)
    x2 = E.meshgen.box_scene_state(X2, seed=1, centre=C3B)
    _same(oracle.ref_cd(fn2, x2, THR, None, None, whd[:1], Em[:1], 0), oracle.ref_cd(fn2, x2, THR, None, None, whd[:1], Em[:1], 0, adapter=True), "second mesh")


# ---- Forces::fill -------------------------------------------------------------------------------------------------------------------
# adapter/Forces_fill_b200.cpp — the body a maintainer puts in place of Forces.cpp:912-930 — compiled against the reference's own Forces.h
# (mesh.hpp, <Eigen/Dense>, <Eigen/Sparse> = oracle/mini_eigen) and linked with the reference's ArcSim mesh code and the same driver as
# libforces_ref.so (oracle/_ref/libadapter_forces.so).  The call goes
#     Mesh (built by ArcSim's own Mesh::add) + Material + Vector3d  ->  Forces::fill (adapter)  ->  eolc::host::flatten / Forces  ->  C ABI  ->  GPU
# and comes back in the members Eigen::VectorXd f, Eigen::SparseMatrix<double> M, MDK, int EoL_cutoff of the reference's class Forces.
# It must equal the reference's own Forces::fill (libforces_ref.so): identical compressed index arrays, values within 1e-10.
from util import assert_close_tol, block_row_scale  # noqa: E402


def _same_fill(ref, got, n_nodes, what):
    assert ref["dof"] == got["dof"] and ref["EoL_cutoff"] == got["EoL_cutoff"], what
    for k in ("M", "MDK"):
        assert np.array_equal(ref[k][0], got[k][0]) and np.array_equal(ref[k][1], got[k][1]), f"{what}: {k} index arrays differ"
        assert_close_tol(got[k][2], ref[k][2], block_row_scale(ref[k][0], ref[k][2], n_nodes), 1e-10, f"{what} {k}")
    assert_close_tol(got["f"], ref["f"], np.abs(ref["f"]).max(), 1e-10, f"{what} f")


MODES = [True, "zero_copy"]    # results copied into the Eigen members / the members' value arrays are the DMA targets (-DEOLC_ADAPTER_ZERO_COPY)


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("gen,n", [("regular2", 24), ("build4", 13), ("regular2", 3)])
def test_adapter_forces_fill_equals_reference(oracle, gen, n, mode):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.drape_state(X, seed=n)
    ref = oracle.ref_forces_fill(fn, x, X)
    got = oracle.ref_forces_fill(fn, x, X, adapter=mode)
    _same_fill(ref, got, X.shape[0], f"{gen}{n}")
    K = got["MDK"]
    import scipy.sparse as sp
    A = sp.csc_matrix((K[2], K[1], K[0]), shape=(got["dof"], got["dof"]))
    assert abs(A - A.T).max() == 0.0          # exactly symmetric, like the reference's mirrored triplets


@pytest.mark.parametrize("mode", MODES)
def test_adapter_forces_fill_steps_on_the_same_objects(oracle, mode):
    """Three fills on the same Mesh / Forces objects: the adapter keeps its plan and page-locked buffers; the second step moves x only
    (M is not recomputed nor copied: the member must still hold it), the third also moves the material coordinates (M changes)."""
    X, fn = E.meshgen.regular2(20)
    rng = np.random.default_rng(9)
    x = E.meshgen.drape_state(X, seed=1)
    x2 = x + 1e-3 * rng.standard_normal(x.shape)
    X3 = X + 2e-4 * rng.standard_normal(X.shape)
    steps = [(x2, None), (x2, X3)]
    refs = oracle.ref_forces_fill(fn, x, X, more_steps=steps)
    gots = oracle.ref_forces_fill(fn, x, X, adapter=mode, more_steps=steps)
    for i, (r, g) in enumerate(zip(refs, gots)):
        _same_fill(r, g, X.shape[0], f"step {i}")
    assert refs[0]["M"][2].tobytes() == refs[1]["M"][2].tobytes() != refs[2]["M"][2].tobytes()
    assert gots[0]["M"][2].tobytes() == gots[1]["M"][2].tobytes() != gots[2]["M"][2].tobytes()


@pytest.mark.parametrize("mode", MODES)
def test_adapter_forces_fill_with_eol_nodes(oracle, mode):
    """EoL nodes: dof = 3N + 2 EoL_Count, Eulerian blocks, EoL_cutoff — through the adapter."""
    X, fn = E.meshgen.regular2(16)
    N = X.shape[0]
    eol = np.full(N, -1, np.int32)
    line = np.arange(1, 15) * 16 + 8
    eol[line] = np.arange(line.size)
    x = E.meshgen.drape_state(X, seed=2)
    ref = oracle.ref_forces_fill(fn, x, X, eol_index=eol)
    got = oracle.ref_forces_fill(fn, x, X, eol_index=eol, adapter=mode)
    assert ref["dof"] == 3 * N + 2 * line.size
    _same_fill(ref, got, N, "EOL line")


@pytest.mark.parametrize("mode", MODES)
def test_adapter_forces_fill_step_time(oracle, capsys, mode):
    """Several steady steps through the adapter on the same objects (the driver moves every node a little before each): values equal
    the reference's own fill of the same states; then the step time on a 512^2 sheet, ArcSim pointer mesh in, Eigen members out.
    Default mode: the values pass through the host layer's page-locked arrays and are copied into the members on the host threads;
    zero-copy mode: the members' value arrays are page-locked in place once and the device -> host copy lands in them."""
    X, fn = E.meshgen.regular2(96)
    x = E.meshgen.drape_state(X, seed=3)
    rng = np.random.default_rng(1)
    steps = [(x + 1e-4 * rng.standard_normal(x.shape), None) for _ in range(4)]
    refs = oracle.ref_forces_fill(fn, x, X, more_steps=steps)
    gots = oracle.ref_forces_fill(fn, x, X, adapter=mode, more_steps=steps)
    for i, (r, g) in enumerate(zip(refs, gots)):
        _same_fill(r, g, X.shape[0], f"step {i}")
    X, fn = E.meshgen.regular2(512)
    first, steady = oracle.ref_forces_step_seconds(fn, E.meshgen.drape_state(X, seed=0), X, steps=5, adapter=mode)
    with capsys.disabled():
        print(f"\n[adapter Forces::fill ({'zero copy' if mode == 'zero_copy' else 'copy'}), regular2 512^2, ArcSim mesh -> Eigen members] first call {first * 1e3:.1f} ms, steady step {steady * 1e3:.2f} ms")
    assert steady < 0.1           # 412 MB per step; 13-20 ms on a B200 box
