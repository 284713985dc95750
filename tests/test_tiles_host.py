"""The "tiles" Forces::fill pipeline WITHOUT a GPU: the plan builder (csrc/forces_plan.h) and the two per-thread phases
(csrc/tile_exec.cuh) are the same source the CUDA kernel compiles; here they are compiled for the host and the threads of a
tile run one after the other (tests/hostmath/hostmath.cpp).  Checked against the oracle to the GPU tolerance (1e-10)."""
import ctypes

import numpy as np
import pytest

import eol_cloth_b200 as E
from util import assert_close_tol, block_row_scale, fan_mesh, strip_mesh

MAT = E.Material.DEFAULT
GRAV = (0.0, 0.0, -9.8)
H = 0.5e-2
ip = ctypes.POINTER(ctypes.c_int32)
dp = ctypes.POINTER(ctypes.c_double)
lp = ctypes.POINTER(ctypes.c_int64)


class HostTiles:
    def __init__(self, L, N, fn, es, X_hint=None, dedup=True, eol_index=None):
        self.L = L
        L.hm_plan_create_eol.restype = ctypes.c_void_p
        L.hm_plan_create_eol.argtypes = [ctypes.c_int32, ctypes.c_int32, ip, ctypes.c_int32, ip, dp, ctypes.c_int, ip, ctypes.c_char_p, ctypes.c_int]
        L.hm_plan_destroy.argtypes = [ctypes.c_void_p]
        L.hm_plan_info.argtypes = [ctypes.c_void_p, lp]
        L.hm_plan_pattern.argtypes = [ctypes.c_void_p, ctypes.c_int, ip, ip]
        L.hm_plan_fill.restype = ctypes.c_int
        L.hm_plan_fill.argtypes = [ctypes.c_void_p, dp, dp, dp, dp, ctypes.c_double, dp, dp, dp]
        L.hm_plan_fill_ex.restype = ctypes.c_int
        L.hm_plan_fill_ex.argtypes = [ctypes.c_void_p, dp, dp, dp, dp, ctypes.c_double, dp, dp, dp, ctypes.c_int]
        fn = np.ascontiguousarray(fn, np.int32).reshape(-1, 3)
        es = np.ascontiguousarray(es, np.int32).reshape(-1, 4)
        err = ctypes.create_string_buffer(256)
        xh = None if X_hint is None else np.ascontiguousarray(X_hint, np.float64)
        eol = None if eol_index is None else np.ascontiguousarray(eol_index, np.int32)
        self.h = L.hm_plan_create_eol(N, len(fn), fn.ctypes.data_as(ip), len(es), es.ctypes.data_as(ip),
                                      None if xh is None else xh.ctypes.data_as(dp), int(dedup), None if eol is None else eol.ctypes.data_as(ip), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        info = np.zeros(17, np.int64)
        L.hm_plan_info(self.h, info.ctypes.data_as(lp))
        self.info = dict(zip("nnzM nnzK n_tiles n_templates elem_evals geo_bytes tmpl_bytes max_scratch max_loc Ei n_runs n_groups pull_rows max_kstage max_mstage dof service_p3".split(), info.tolist()))
        self.N = N
        self.dof = self.info["dof"]

    def pattern(self, which):
        nnz = self.info["nnzK" if which else "nnzM"]
        o, i = np.zeros(self.dof + 1, np.int32), np.zeros(max(nnz, 1), np.int32)
        self.L.hm_plan_pattern(self.h, which, o.ctypes.data_as(ip), i.ctypes.data_as(ip))
        return o, i[:nnz]

    def fill(self, x, X, mat=MAT, grav=GRAV, h=H, phases=(0, 0, 0)):
        """phases: 16-byte phase (0 / 1 doubles) of the f, M, MDK output pointers — the bulk copy-out must cope with all of them."""
        x = np.ascontiguousarray(x, np.float64); X = np.ascontiguousarray(X, np.float64)

        def out(n, ph):
            buf = np.full(n + 3, np.nan)
            off = ((-buf.ctypes.data // 8) % 2 + ph) % 2 if buf.ctypes.data % 16 in (0, 8) else 0
            return buf, buf[off:off + n]
        (fb, f), (Mb, Mv), (Kb, Kv) = out(self.dof, phases[0]), out(self.info["nnzM"], phases[1]), out(self.info["nnzK"], phases[2])
        for a, ph in ((f, phases[0]), (Mv, phases[1]), (Kv, phases[2])):
            assert (a.ctypes.data // 8) % 2 == ph
        m = np.array(mat, np.float64); g = np.array(grav, np.float64)
        rc = self.L.hm_plan_fill(self.h, x.ctypes.data_as(dp), X.ctypes.data_as(dp), m.ctypes.data_as(dp), g.ctypes.data_as(dp), h,
                                 f.ctypes.data_as(dp), Mv.ctypes.data_as(dp), Kv.ctypes.data_as(dp))
        assert rc == 0, "a bulk copy was issued with a misaligned address or size"
        for b, a in ((fb, f), (Mb, Mv), (Kb, Kv)):   # nothing written outside the arrays
            assert np.isnan(b).sum() == b.size - a.size + np.isnan(a).sum()
        return f, Mv, Kv

    def close(self):
        self.L.hm_plan_destroy(self.h)


def test_tiles_host_m_unchanged_mode(hostmath):
    """EOLC_FILL_M_UNCHANGED: f and MDK bit-identical to the full fill, not one M value written."""
    for gen, n in (("regular2", 19), ("build4", 8)):
        X, fn = getattr(E.meshgen, gen)(n)
        es = E.meshgen.edge_stencils(X.shape[0], fn)
        x = E.meshgen.drape_state(X, seed=n)
        T = HostTiles(hostmath, X.shape[0], fn, es, X, True)
        f, Mv, Kv = T.fill(x, X)
        f2, M2, K2 = np.full_like(f, np.nan), np.full(Mv.size + 1, np.nan)[:Mv.size], np.full_like(Kv, np.nan)
        m = np.array(MAT, np.float64); g = np.array(GRAV, np.float64)
        x = np.ascontiguousarray(x); Xc = np.ascontiguousarray(X)
        rc = hostmath.hm_plan_fill_ex(T.h, x.ctypes.data_as(dp), Xc.ctypes.data_as(dp), m.ctypes.data_as(dp), g.ctypes.data_as(dp), H,
                                      f2.ctypes.data_as(dp), M2.ctypes.data_as(dp), K2.ctypes.data_as(dp), 1)
        if (f2.ctypes.data // 8) % 2 == 0 and (K2.ctypes.data // 8) % 2 == 0:
            assert rc == 0
        assert f2.tobytes() == f.tobytes() and K2.tobytes() == Kv.tobytes()
        assert np.isnan(M2).all()
        T.close()


def _check(T, fn, es, x, X, oracle, what, mat=MAT, grav=GRAV, h=H, phases=(0, 0, 0), eol_index=None):
    f, Mv, Kv = T.fill(x, X, mat, grav, h, phases)
    assert not np.isnan(f).any() and not np.isnan(Mv).any() and not np.isnan(Kv).any(), what + ": an output slot was never written"
    ref = oracle.forces_fill(fn, es, x, X, tuple(mat), grav, h, eol_index=eol_index)
    N = x.shape[0]
    assert_close_tol(f, ref["f"], max(np.abs(ref["f"]).max(), 1e-300), 1e-10, what + " f")
    for name, which, got in (("M", 0, Mv), ("MDK", 1, Kv)):
        o, i, v = ref[name]
        po, pi = T.pattern(which)
        assert np.array_equal(po, o) and np.array_equal(pi, i), what + f" {name} pattern"
        assert_close_tol(got, v, block_row_scale(o, v, N), 1e-10, what + f" {name} values")
    return f, Mv, Kv


@pytest.mark.parametrize("gen,n,hint,dedup", [("regular2", 2, True, True), ("regular2", 3, True, True), ("build4", 3, True, False),
                                              ("regular2", 17, True, True), ("build4", 9, False, True), ("regular2", 40, True, True),
                                              ("regular2", 33, False, False)])
def test_tiles_host_matches_oracle(oracle, hostmath, gen, n, hint, dedup):
    X, fn = getattr(E.meshgen, gen)(n)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=n)
    T = HostTiles(hostmath, X.shape[0], fn, es, X if hint else None, dedup)
    _check(T, fn, es, x, X, oracle, f"{gen}{n}")
    _check(T, fn, es, x, X, oracle, f"{gen}{n} odd phases", phases=(1, 1, 1))
    _check(T, fn, es, x, X, oracle, f"{gen}{n} mixed phases", phases=(0, 1, 0))
    assert T.info["elem_evals"] >= len(fn) + T.info["Ei"]
    T.close()


def test_tiles_host_dedup_and_tiling_quality(hostmath):
    """A 128x128 regular sheet: 8x4-node tiles, interior tiles share one template, halo re-evaluation stays below 1.6x."""
    X, fn = E.meshgen.regular2(128)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    T = HostTiles(hostmath, X.shape[0], fn, es, X, True)
    i = T.info
    assert i["n_tiles"] == 128 * 128 // 32, i
    assert i["n_templates"] < 60, i
    assert i["elem_evals"] < 1.6 * (len(fn) + i["Ei"]), i
    assert i["max_scratch"] * 8 <= 112 * 1024
    T2 = HostTiles(hostmath, X.shape[0], fn, es, X, False)
    assert T2.info["n_templates"] == T2.info["n_tiles"]
    x = E.meshgen.drape_state(X, seed=1)
    a, b, c = T.fill(x, X), T2.fill(x, X), T.fill(x, X)
    for u, v, w in zip(a, b, c):
        assert u.tobytes() == w.tobytes()      # same plan, same bits
        # shared templates get a longer bank-layout search (forces_plan.h: optimize_template), which may reorder the additions
        # of a block: same values to rounding
        assert np.allclose(u, v, rtol=0, atol=1e-13 * np.abs(u).max())
    T.close(); T2.close()


def test_tiles_host_shuffled_isolated_material(oracle, hostmath):
    X, fn = E.meshgen.regular2(9)
    rng = np.random.default_rng(11)
    N = X.shape[0] + 1
    perm = rng.permutation(N)
    Xn = np.zeros((N, 2)); Xn[perm[:-1]] = X; Xn[perm[-1]] = (5.0, 5.0)
    fnn = perm[fn].astype(np.int32)
    fnn = fnn[rng.permutation(len(fnn))]
    es = E.meshgen.edge_stencils(N, fnn)
    x = E.meshgen.drape_state(Xn, seed=2)
    T = HostTiles(hostmath, N, fnn, es, Xn, True)
    mat = E.Material(0.2, 1000.0, 0.3, 1e-3, 0.0, 0.7)
    f, _, _ = _check(T, fnn, es, x * 1.3, Xn, oracle, "shuffled", mat, (0.1, -0.2, -9.8), 1e-2)
    assert np.all(f[3 * perm[-1]:3 * perm[-1] + 3] == 0)
    T.close()


@pytest.mark.parametrize("mesh", [fan_mesh(6), fan_mesh(40), fan_mesh(100), strip_mesh(50)], ids=["fan6", "fan40", "fan100", "strip50"])
def test_tiles_host_high_valence_and_strips(oracle, hostmath, mesh):
    T = HostTiles(hostmath, mesh["x"].shape[0], mesh["face_nodes"], mesh["edge_stencil"], mesh["X"], True)
    _check(T, mesh["face_nodes"], mesh["edge_stencil"], mesh["x"], mesh["X"], oracle, "valence")
    T.close()


def test_tiles_host_rejects_nonmanifold(hostmath):
    # two faces glued along an edge AND sharing the opposite vertex: the bending stencil repeats a node
    fn = np.array([[0, 1, 2], [1, 0, 2]], np.int32)
    es = np.array([[0, 1, 2, 2]], np.int32)
    with pytest.raises(RuntimeError):
        HostTiles(hostmath, 3, fn, es, None, True)


def test_tiles_host_plan_independent_of_worker_count(hostmath, monkeypatch):
    """forces_plan.h builds the tiles on several host threads and merges the ranges in tile order: the plan (geometry blobs,
    templates, template order) must be the same bytes for 1, 3 and 8 workers, with and without template sharing."""
    hostmath.hm_plan_hash.restype = ctypes.c_uint64
    hostmath.hm_plan_hash.argtypes = [ctypes.c_void_p]
    X, fn = E.meshgen.regular2(128)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    rng = np.random.default_rng(5)
    perm = rng.permutation(X.shape[0]).astype(np.int32)     # irregular numbering: no template sharing to speak of
    fn_p = perm[fn]
    es_p = np.where(es >= 0, perm[np.maximum(es, 0)], -1).astype(np.int32)
    X_p = np.empty_like(X); X_p[perm] = X
    for args in ((X.shape[0], fn, es, X, True), (X.shape[0], fn_p, es_p, X_p, True), (X.shape[0], fn, es, None, False)):
        hashes, infos = [], []
        for nt in ("1", "3", "8"):
            monkeypatch.setenv("EOLC_PLAN_THREADS", nt)
            T = HostTiles(hostmath, *args)
            hashes.append(hostmath.hm_plan_hash(T.h)); infos.append(T.info)
            T.close()
        assert hashes[0] == hashes[1] == hashes[2], (hashes, infos)
        assert infos[0] == infos[1] == infos[2]


def _eol_line(n, j0):
    """EoL nodes along the grid line j = j0 of an n x n regular2 sheet (the cloth crossing a box edge), indices in a shuffled order."""
    eol = np.full(n * n, -1, np.int32)
    line = np.arange(1, n - 1) * n + j0
    eol[line] = np.random.default_rng(n).permutation(line.size)
    return eol


@pytest.mark.parametrize("gen,n,eol_nodes,dedup", [("regular2", 5, (12, 6, 18, 7), True), ("regular2", 4, (5,), True), ("build4", 3, (4, 9, 10, 1), False),
                                                   ("regular2", 3, tuple(range(9)), True), ("regular2", 24, "line", True), ("regular2", 33, "line", False),
                                                   ("build4", 9, "corner+isolated", True)])
def test_tiles_host_eol_matches_oracle(oracle, hostmath, gen, n, eol_nodes, dedup):
    """EOL branch (forces_eol.h): the tiles plan with moved row destinations + the EOL element records + the gather, emulated on the host,
    against the oracle's restatement of fillEOL* (f, patterns bit-exact, values to 1e-10)."""
    X, fn = getattr(E.meshgen, gen)(n)
    N = X.shape[0]
    es = E.meshgen.edge_stencils(N, fn)
    x = E.meshgen.drape_state(X, seed=n)
    if eol_nodes == "line":
        eol = _eol_line(n, n // 2)
    elif eol_nodes == "corner+isolated":
        eol = np.full(N, -1, np.int32)
        eol[0] = 2; eol[N - 1] = 0; eol[n * n // 2] = 1          # a corner node, a cell-centre node, an interior grid node
    else:
        eol = np.full(N, -1, np.int32)
        for k, a in enumerate(eol_nodes):
            eol[a] = k
    T = HostTiles(hostmath, N, fn, es, X, dedup, eol_index=eol)
    assert T.dof == 3 * N + 2 * (int(eol.max()) + 1)
    _check(T, fn, es, x, X, oracle, f"eol {gen}{n}", eol_index=eol)
    _check(T, fn, es, x, X, oracle, f"eol {gen}{n} odd phases", phases=(1, 1, 1), eol_index=eol)
    _check(T, fn, es, x, X, oracle, f"eol {gen}{n} mixed phases", phases=(1, 0, 1), eol_index=eol)
    T.close()


def test_edge_force_matches_reference(oracle, hostmath):
    """The bending force the EOL branch needs (the Lagrangian one drops it): forces_eol.h edge_force vs the reference's ComputeBending f."""
    hostmath.hostmath_edge_force.argtypes = [dp] * 8 + [ctypes.c_double, dp]
    rng = np.random.default_rng(8)
    for trial in range(50):
        X = np.array([[0, 0], [1, 0], [0.2, 0.9], [0.7, -0.8]]) + 0.1 * rng.standard_normal((4, 2))
        x = np.c_[X, np.zeros(4)] + 0.2 * rng.standard_normal((4, 3))
        _, f, _ = oracle.compute_bending(*x, *X, 1e-2)
        got = np.zeros(12)
        hostmath.hostmath_edge_force(*[np.ascontiguousarray(v).ctypes.data_as(dp) for v in (*x, *X)], 1e-2, got.ctypes.data_as(dp))
        assert np.abs(got - f).max() <= 1e-12 * np.abs(f).max(), trial


def test_tiles_host_eol_shuffled_numbering_and_faceless_eol_node(oracle, hostmath):
    """Remeshed-like node numbering (runs of consecutive nodes are short, coupled and uncoupled nodes interleave) and an EoL node no
    face refers to: its two Eulerian rows stay empty and its f entries are written as zeros."""
    X, fn = E.meshgen.regular2(11)
    rng = np.random.default_rng(3)
    N = X.shape[0] + 1
    perm = rng.permutation(N).astype(np.int32)
    Xp = np.zeros((N, 2)); Xp[perm[:-1]] = X; Xp[perm[-1]] = (0.5, 0.5)
    fnp = perm[fn]
    es = E.meshgen.edge_stencils(N, fnp)
    x = np.c_[Xp, 0.05 * np.sin(5 * Xp[:, 0]) * np.cos(3 * Xp[:, 1])] + 1e-3 * rng.standard_normal((N, 3))
    eol = np.full(N, -1, np.int32)
    chosen = perm[[5 * 11 + j for j in range(2, 9)]]          # a grid line of the original sheet
    eol[chosen] = rng.permutation(len(chosen)) + 1
    eol[perm[-1]] = 0                                         # the faceless node is EoL index 0
    T = HostTiles(hostmath, N, fnp, es, Xp, True, eol_index=eol)
    f, Mv, Kv = _check(T, fnp, es, x, Xp, oracle, "eol shuffled", eol_index=eol)
    assert f[3 * N] == 0.0 and f[3 * N + 1] == 0.0
    o, _ = T.pattern(1)
    assert o[3 * N] == o[3 * N + 1] == o[3 * N + 2]
    T.close()


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_tiles_host_eol_random_triangulations(oracle, hostmath, seed):
    """Irregular meshes (Delaunay triangulations of random points: valences 3..10, no structure for the templates to share) with a random
    third of the nodes EoL, so that faces and stencils with 1, 2, 3 and 4 EoL vertices all occur."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(100 + seed)
    n = 60 + 25 * seed
    X = rng.uniform(0, 1, (n, 2))
    tri = Delaunay(X).simplices.astype(np.int32)
    # counter-clockwise faces
    a, b, c = X[tri[:, 0]], X[tri[:, 1]], X[tri[:, 2]]
    flip = ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0])) < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]
    area = np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
    tri = tri[area > 1e-4]                       # slivers on the hull would only test the conditioning of DX^-1
    es = E.meshgen.edge_stencils(n, tri)
    x = np.c_[X, 0.1 * np.sin(4 * X[:, 0]) * np.cos(3 * X[:, 1])] + 2e-3 * rng.standard_normal((n, 3))
    eol = np.full(n, -1, np.int32)
    chosen = rng.choice(n, n // 3, replace=False)
    eol[chosen] = rng.permutation(len(chosen))
    T = HostTiles(hostmath, n, tri, es, X, True, eol_index=eol)
    _check(T, tri, es, x, X, oracle, f"eol delaunay {seed}", eol_index=eol)
    _check(T, tri, es, x, X, oracle, f"eol delaunay {seed} odd phases", phases=(1, 1, 1), eol_index=eol)
    T.close()


def test_phase3_placement_follows_the_tile_population(hostmath):
    """forces_plan.h mostly_full_tiles(): phase 3 goes to the kernel's service warps for sheets whose tiles are nearly all interior ones
    (measured -1.7 % at 1024^2) and stays on the compute warps for the 64x64 scenes of the ensemble (+2.5 % there otherwise)."""
    for n, want in ((64, 0), (256, 1)):
        X, fn = E.meshgen.regular2(n)
        es = E.meshgen.edge_stencils(X.shape[0], fn)
        T = HostTiles(hostmath, X.shape[0], fn, es, X, True)
        assert T.info["service_p3"] == want, (n, T.info)
        T.close()
