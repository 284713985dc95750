"""Pins the CPU oracle: reference object code vs SURVEY §8c known answers, RNG / raytri KATs, FD identities,
pattern sizes, and the committed golden fixtures (tests/golden, made by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import eol_cloth_b200 as E

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

x0 = (0, 0, 0); x1 = (1.1, 0.05, 0.1); x2 = (-0.02, 0.9, -0.05); x3 = (1, -1, 0.2)
X0 = (0, 0); X1 = (1, 0); X2 = (0, 1); X3 = (1, -1)


def test_known_answers_membrane(oracle):
    W, f, K = oracle.compute_membrane(x0, x1, x2, X0, X1, X2, 50, 0.01, [1, 0, 0, 1, 0, 0], [1, 0, 0, 1])
    assert W == 0.28341584158415856
    assert f[0] == 1.9801980198019824
    assert K[0, 0] == 49.757526773085473
    assert K[0, 8] == 0 and np.signbit(K[0, 8])
    assert abs(np.abs(K).sum() - 400.08082440897147) < 1e-11
    assert np.array_equal(K, K.T)


def test_known_answers_bending(oracle):
    W, f, K = oracle.compute_bending(x0, x1, x2, x3, X0, X1, X2, X3, 1e-5)
    assert W == 1.9012987629236648e-08
    assert f[0] == -8.5073508439304069e-08 and f[11] == -7.1163691863669505e-07
    assert K[0, 0] == 3.3642796219437894e-07 and K[0, 11] == 1.5876586965473499e-06
    assert abs(np.abs(K).sum() - 0.00034194309149763314) < 1e-17
    assert np.array_equal(K, K.T)


def test_known_answers_inertial(oracle):
    W, f, M = oracle.compute_inertial(x0, x1, x2, X0, X1, X2, (0, 0, -9.8), 0.05)
    assert W == 0.0040833333333333372
    assert f[2] == -0.081666666666666679
    assert M[0, 0] == 0.0041666666666666666 and M[0, 3] == 0.0020833333333333333 and M[0, 1] == 0


def test_fd_identities(oracle):
    """f = -dW/dx and K = -df/dx for the reference kernels (central differences)."""
    rng = np.random.default_rng(3)
    X = np.array([[0, 0], [1, 0], [0.2, 0.9], [0.7, -0.8]]) + 0.05 * rng.standard_normal((4, 2))
    x = np.c_[X, np.zeros(4)] + 0.05 * rng.standard_normal((4, 3))
    h = 1e-6
    W, f, K = oracle.compute_bending(*x, *X, 1e-2)
    for i in range(12):
        xp = x.copy().reshape(-1); xm = xp.copy(); xp[i] += h; xm[i] -= h
        Wp, fp, _ = oracle.compute_bending(*xp.reshape(4, 3), *X, 1e-2)
        Wm, fm, _ = oracle.compute_bending(*xm.reshape(4, 3), *X, 1e-2)
        assert abs(-(Wp - Wm) / (2 * h) - f[i]) < 1e-7 * max(1, np.abs(f).max())
        assert np.abs(-(fp - fm) / (2 * h) - K[:, i]).max() < 1e-6 * np.abs(K).max()
    P, Q = oracle.face_frame(*x[:3], *X[:3])
    W, f, K = oracle.compute_membrane(*x[:3], *X[:3], 50.0, 0.3, P, Q)
    for i in range(9):
        xp = x[:3].copy().reshape(-1); xm = xp.copy(); xp[i] += h; xm[i] -= h
        Wp, fp, _ = oracle.compute_membrane(*xp.reshape(3, 3), *X[:3], 50.0, 0.3, P, Q)   # frozen frame
        Wm, fm, _ = oracle.compute_membrane(*xm.reshape(3, 3), *X[:3], 50.0, 0.3, P, Q)
        assert abs(-(Wp - Wm) / (2 * h) - f[i]) < 1e-6 * max(1, np.abs(f).max())
        assert np.abs(-(fp - fm) / (2 * h) - K[:, i]).max() < 1e-6 * np.abs(K).max()


def test_rng_known_answer(oracle):
    want = [0.99436961646053112, 0.86511472273633094, -0.74375110445538795, 0.99808103093054723,
            -0.52782204740366157, -0.20683854767478138]
    got = oracle.rng_raw(6)
    assert got.tolist() == want
    assert got[:1].view(np.uint64)[0] == 0x3FEFD1E03ADAB07E
    p = oracle.perturbation(2, 5e-3)
    assert p[0, 0] == want[0] * 5e-3 * 1e-3


def test_raytri_borders(oracle):
    """ZERO = -EPSILON makes the borders inset by 1e-6 (scaled by det), SURVEY §8a row 16."""
    v0, v1, v2 = (0, 0, 0), (1, 0, 0), (0, 1, 0)
    d = (0, 0, -1)
    assert oracle.raytri((0.25, 0.25, 1), d, v0, v1, v2)[0] == 1
    assert oracle.raytri((0.5, 0.0, 1), d, v0, v1, v2)[0] == 0          # on an edge
    assert oracle.raytri((0.5, 5e-7, 1), d, v0, v1, v2)[0] == 0         # 5e-7 inside
    assert oracle.raytri((0.5, 2e-6, 1), d, v0, v1, v2)[0] == 1         # 2e-6 inside
    hit, tuv = oracle.raytri((0.25, 0.25, 1), d, v0, v1, v2)
    assert tuv[0] == 1.0


@pytest.mark.parametrize("gen,n,nnzM,nnzK", [("regular2", 3, 369, 513), ("build4", 3, 621, 837)])
def test_pattern_sizes_small(oracle, gen, n, nnzM, nnzK):
    X, fn = getattr(E.meshgen, gen)(n)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    r = oracle.forces_fill(fn, es, E.meshgen.drape_state(X), X)
    assert r["M"][2].size == nnzM and r["MDK"][2].size == nnzK
    # Eigen-compressed invariants: sorted inner indices, exact symmetry
    for name in ("M", "MDK"):
        outer, inner, vals = r[name]
        for c in range(outer.size - 1):
            seg = inner[outer[c]:outer[c + 1]]
            assert np.all(np.diff(seg) > 0)


def test_survey_counts_64(oracle):
    X, fn = E.meshgen.regular2(64)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    assert X.shape[0] == 4096 and fn.shape[0] == 7938 and int((es[:, 3] >= 0).sum()) == 11781
    r = oracle.forces_fill(fn, es, E.meshgen.drape_state(X), X)
    assert r["M"][2].size == 253458 and r["MDK"][2].size == 465516


def _close_to_golden(r, g, n_nodes):
    """The fixtures are written by THE REFERENCE'S OWN compiled Forces::fill (oracle/_ref/libforces_ref.so, tests/golden/make_golden.py);
    the Eigen-free restatement reproduces them: identical index arrays, M bit for bit, f / MDK to 1e-13 of the block-row scale (they
    differ in last bits only where Eigen's operators order a sum differently; tests/test_forces_ref_pin.py has the full comparison)."""
    from util import assert_close_tol, block_row_scale
    assert str(g["source"]) == "libforces_ref"
    assert_close_tol(r["f"], g["f"], np.abs(g["f"]).max(), 1e-13, "f")
    for k, nm in (("M", "M"), ("K", "MDK")):
        assert np.array_equal(r[nm][0], g[k + "_outer"]) and np.array_equal(r[nm][1], g[k + "_inner"])
        assert_close_tol(r[nm][2], g[k + "_vals"], block_row_scale(g[k + "_outer"], g[k + "_vals"], n_nodes), 1e-13, nm)
    assert np.array_equal(r["M"][2], g["M_vals"])


@pytest.mark.parametrize("name,gen,n,seed", [("forces_regular2_n12", "regular2", 12, 0), ("forces_build4_n7", "build4", 7, 1)])
def test_golden_forces(oracle, name, gen, n, seed):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    X, fn = getattr(E.meshgen, gen)(n)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    r = oracle.forces_fill(fn, es, E.meshgen.drape_state(X, seed=seed), X)
    _close_to_golden(r, g, X.shape[0])


def _eol_line(n):
    eol = np.full(n * n, -1, np.int32)
    line = np.arange(1, n - 1) * n + n // 2
    eol[line] = np.arange(line.size)
    return eol


def test_golden_forces_eol_and_normals(oracle):
    g = np.load(os.path.join(GOLD, "forces_eol_regular2_n12.npz"))
    X, fn = E.meshgen.regular2(12)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    r = oracle.forces_fill(fn, es, E.meshgen.drape_state(X, seed=2), X, eol_index=_eol_line(12))
    assert r["dof"] == 3 * 144 + 2 * 10
    _close_to_golden(r, g, X.shape[0])
    g = np.load(os.path.join(GOLD, "normals_build4_n7.npz"))
    X, fn = E.meshgen.build4(7)
    fa, na = oracle.mesh_normals(fn, E.meshgen.drape_state(X, seed=1))
    assert np.array_equal(fa, g["face_n"]) and np.array_equal(na, g["node_n"])


def _cd_inputs(gen, n, centre, seed, points):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
    pxyz = pn = None
    if points:
        pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3])
        pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]])
    return X, fn, x, pxyz, pn


CD_CASES = [("cd_regular2_n24", "regular2", 24, tuple(E.meshgen.BOX_CENTRE), 0, False),
            ("cd_build4_n16_corner", "build4", 16, (0.9175, -0.25, -0.549), 1, True)]


@pytest.mark.parametrize("name,gen,n,centre,seed,points", CD_CASES)
def test_golden_cd(oracle, name, gen, n, centre, seed, points):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    X, fn, x, pxyz, pn = _cd_inputs(gen, n, centre, seed, points)
    for key, flag, remap in (("cd", 1, 1), ("cd2", 0, 0)):
        got = oracle.cd(fn, x, E.meshgen.BOX_THRESHOLD, pxyz, pn, E.meshgen.BOX_WHD[None], E.meshgen.box_frame(centre)[None], flag, remap)
        assert got.tobytes() == g[key].tobytes()
    types = set(zip(g["cd"]["count1"].tolist(), g["cd"]["count2"].tolist()))
    if points:
        assert types == {(1, 3), (2, 2), (3, 1)}     # all three contact types are covered


def test_cd_edge_table_matches_createEdges_order(oracle):
    X, fn = E.meshgen.regular2(5)
    tab, _ = oracle.cd_edges(fn, np.c_[X, np.zeros(len(X))])
    # sorted by (max, min) of the edge's endpoints
    key = np.maximum(tab[:, 0], tab[:, 1]).astype(np.int64) * 10**6 + np.minimum(tab[:, 0], tab[:, 1])
    assert np.all(np.diff(key) > 0)
    assert tab.shape[0] == 3 * 16 + 2 * 4


def test_eigen_cg_restatement_against_direct_solve(oracle):
    """oracle.eigen_cg (restated Eigen 3.3 CG, the reference's collision-free solver branch) against a sparse direct solve of the
    assembled system; cloth_rhs against dense arithmetic."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl
    import eol_cloth_b200 as E
    X, fn = E.meshgen.regular2(10)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=3)
    ref = oracle.forces_fill(fn, es, x, X)
    dof = ref["f"].size
    rng = np.random.default_rng(0)
    v = 0.1 * rng.standard_normal(dof)
    h = 0.5e-2
    b = oracle.cloth_rhs(ref["M"], ref["f"], v, h)
    o, i, vals = ref["M"]
    Md = sp.csc_matrix((vals, i, o), shape=(dof, dof)).toarray()
    assert np.allclose(b, -(Md @ v + h * ref["f"]), rtol=1e-13, atol=1e-18)
    o, i, vals = ref["MDK"]
    K = sp.csc_matrix((vals, i, o), shape=(dof, dof))
    sol, it, res = oracle.eigen_cg(ref["MDK"], b, tol=1e-13)
    direct = spl.spsolve(K.tocsc(), -b)
    assert res < 1e-13 and 0 < it <= 2 * dof
    assert np.abs(sol - direct).max() <= 1e-9 * np.abs(direct).max()
    assert oracle.eigen_cg(ref["MDK"], np.zeros(dof))[1] == 0


def test_threaded_timing_variant_is_bit_identical(oracle):
    """bench.py's CPU legs run the oracle's element loops on several threads: same triplet sequence, same sums."""
    X, fn = E.meshgen.regular2(40)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=4)
    eol = np.full(X.shape[0], -1, np.int32)
    eol[np.arange(1, 39) * 40 + 20] = np.arange(38)
    for e in (None, eol):
        a = oracle.forces_fill(fn, es, x, X, eol_index=e)
        b = oracle.forces_fill(fn, es, x, X, eol_index=e, threads=5)
        assert a["f"].tobytes() == b["f"].tobytes()
        for k in ("M", "MDK"):
            for u, v in zip(a[k], b[k]):
                assert u.tobytes() == v.tobytes()


def test_product_free_generators_equal_the_products(oracle):
    """bench.py --impl reference builds its inputs without loading the product: oracle.sheet_regular2 / arcsim_edge_stencils /
    drape_state must be the product's meshgen.regular2 / edge_stencils (eolc_mesh_edge_stencils) / drape_state, and a strip sample must
    be a prefix of the full sheet (coordinates, numbering, faces, state)."""
    import eol_cloth_b200 as E
    for gen, n in (("regular2", 3), ("regular2", 17), ("build4", 9), ("regular2", 64)):
        X, fn = getattr(E.meshgen, gen)(n)
        assert np.array_equal(oracle.arcsim_edge_stencils(len(X), fn), E.meshgen.edge_stencils(len(X), fn)), (gen, n)
    X, fn = E.meshgen.regular2(33)
    X2, fn2 = oracle.sheet_regular2(33)
    assert np.array_equal(X, X2) and np.array_equal(fn, fn2)
    assert np.array_equal(E.meshgen.drape_state(X, seed=4), oracle.drape_state(X2, seed=4))
    Xs, fs = oracle.sheet_regular2(33, rows=9)
    assert np.array_equal(Xs, X[:9 * 33]) and np.array_equal(fs, fn[:2 * 8 * 32])
    assert np.array_equal(oracle.drape_state(Xs, seed=4, n_total=33 * 33), E.meshgen.drape_state(X, seed=4)[:9 * 33])
    perm = np.random.default_rng(0).permutation(len(fn))          # any face order: edges appear in first-use order
    assert np.array_equal(oracle.arcsim_edge_stencils(len(X), fn[perm]), E.meshgen.edge_stencils(len(X), fn[perm]))
