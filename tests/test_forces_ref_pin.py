"""Pins the Forces half of the oracle to THE REFERENCE'S OWN CODE.

oracle/_ref/libforces_ref.so is /root/reference/src/Forces.cpp, UtilEOL.cpp, conversions.cpp, Compute{Membrane,Bending,Inertial}.cpp and
ArcSim's mesh.cpp / geometry.cpp / util.cpp / vectors.cpp / transformation.cpp compiled UNMODIFIED (oracle/Makefile) against
oracle/mini_eigen, a functional stand-in for the Eigen subset they use.  So Forces::fill (Forces.cpp:912-930) with faceBasedF /
edgeBasedF, their EOL branches, poldec and deform_grad runs here as the reference wrote it, on a mesh built the way Cloth::build builds
it (Mesh::add(Face*) creates the edges).  What this file holds to it:

  * oracle/forces_ref.cpp, the Eigen-free restatement the GPU is compared with at sizes the reference needs minutes / tens of GB for:
    identical CSR index arrays, values within 1e-13 of the block-row scale (they differ in the last bits only where Eigen's operators
    — restated in mini_eigen — order a sum differently; the parity budget of the GPU is 1e-10);
  * the edge order: ArcSim's own mesh.edges == oracle.arcsim_edge_stencils == the library's eolc_mesh_edge_stencils;
  * face and node normals of compute_ws_data == oracle.mesh_normals, bit for bit;
  * SURVEY §8c's hand-enumerated pattern sizes, now counted by the reference itself.

No GPU needed.  The prebuilt library travels to the GPU box, where tests/test_forces_gpu.py compares the CUDA path with it directly."""
import numpy as np
import pytest

import eol_cloth_b200 as E
from util import assert_close_tol, block_row_scale, fan_mesh, strip_mesh

TOL = 1e-13


def _compare(ref, got, n_nodes, what, tol=TOL):
    assert ref["dof"] == got["dof"], what
    for k in ("M", "MDK"):
        assert np.array_equal(ref[k][0], got[k][0]), f"{what}: {k} outer index differs"
        assert np.array_equal(ref[k][1], got[k][1]), f"{what}: {k} inner index differs"
        assert_close_tol(got[k][2], ref[k][2], block_row_scale(ref[k][0], ref[k][2], n_nodes), tol, f"{what} {k}")
    assert_close_tol(got["f"], ref["f"], np.abs(ref["f"]).max(), tol, f"{what} f")


def _eol_some(N, count, seed):
    eol = np.full(N, -1, np.int32)
    sel = np.random.default_rng(seed).permutation(N)[:count]
    eol[sel] = np.random.default_rng(seed + 1).permutation(count)   # EoL_index in an order unrelated to the node order
    return eol


def _case(name):
    if name.startswith("regular2") or name.startswith("build4"):
        gen, n = name.split("_")[0], int(name.split("_")[1])
        X, fn = getattr(E.meshgen, gen)(n)
        return X, fn, E.meshgen.drape_state(X, seed=n)
    if name.startswith("fan"):
        m = fan_mesh(int(name.split("_")[1]))
        return m["X"], m["face_nodes"], m["x"]
    if name.startswith("strip"):
        m = strip_mesh(int(name.split("_")[1]))
        return m["X"], m["face_nodes"], m["x"]
    if name.startswith("delaunay"):
        from scipy.spatial import Delaunay
        rng = np.random.default_rng(7)
        X = rng.random((int(name.split("_")[1]), 2))
        tri = Delaunay(X).simplices.astype(np.int32)
        a, b, c = X[tri[:, 0]], X[tri[:, 1]], X[tri[:, 2]]
        flip = (b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]) < 0
        tri[flip] = tri[flip][:, [0, 2, 1]]
        area2 = np.abs((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (b[:, 1] - a[:, 1]) * (c[:, 0] - a[:, 0]))
        tri = np.ascontiguousarray(tri[area2 > 1e-6])            # hull slivers: rest area ~ 0 is a division by ~ 0 in every implementation
        keep = np.unique(tri)
        remap = np.full(X.shape[0], -1, np.int32); remap[keep] = np.arange(keep.size, dtype=np.int32)
        X = np.ascontiguousarray(X[keep]); tri = remap[tri]
        x = np.c_[X, 0.05 * np.sin(5 * X[:, 0]) * np.cos(4 * X[:, 1])] + 1e-3 * rng.standard_normal((X.shape[0], 3))
        return X, tri, x
    raise ValueError(name)


@pytest.mark.parametrize("name", ["regular2_3", "regular2_12", "regular2_41", "build4_3", "build4_7", "build4_20", "fan_5", "fan_17", "strip_9",
                                  "delaunay_300"])
def test_restatement_equals_the_reference_fill(oracle, name):
    X, fn, x = _case(name)
    ref = oracle.ref_forces_fill(fn, x, X)
    es = oracle.arcsim_edge_stencils(X.shape[0], fn)
    assert np.array_equal(ref["edge_stencil"], es), "mesh.edges of ArcSim's Mesh::add(Face*) vs the restated edge order"
    assert np.array_equal(E.meshgen.edge_stencils(X.shape[0], fn), es), "eolc_mesh_edge_stencils"
    assert ref["EoL_cutoff"] == 3 * X.shape[0] == ref["dof"]
    _compare(ref, oracle.forces_fill(fn, es, x, X), X.shape[0], name)
    # the reference's M is bit for bit the restatement's (sums of rho * 2A / 12 and / 24 in face order; no Eigen arithmetic involved)
    assert ref["M"][2].tobytes() == oracle.forces_fill(fn, es, x, X)["M"][2].tobytes()


@pytest.mark.parametrize("name,count", [("regular2_12", 5), ("regular2_6", 36), ("build4_9", 30), ("delaunay_120", 25), ("fan_9", 1), ("fan_9", 10)])
def test_restatement_equals_the_reference_fill_with_eol_nodes(oracle, name, count):
    """The EOL branch (Forces.cpp:177-329, 399-497, 580-683, 746-883; deform_grad UtilEOL.cpp:13-28) as the reference wrote it:
    dof = 3N + 2 EoL_Count, Eulerian rows / columns at 3N + 2 EoL_index, the bending force that only enters f here."""
    X, fn, x = _case(name)
    N = X.shape[0]
    eol = _eol_some(N, min(count, N), 11) if count < N else np.arange(N, dtype=np.int32)
    if name == "fan_9" and count == 1:
        eol = np.full(N, -1, np.int32); eol[0] = 0                      # the hub: every element of the mesh is an EOL element
    ref = oracle.ref_forces_fill(fn, x, X, eol_index=eol)
    assert ref["dof"] == 3 * N + 2 * int((eol >= 0).sum()) and ref["EoL_cutoff"] == 3 * N
    got = oracle.forces_fill(fn, oracle.arcsim_edge_stencils(N, fn), x, X, eol_index=eol)
    _compare(ref, got, N, f"{name} EOL x{count}")
    assert np.abs(ref["f"][3 * N:]).max() > 0


def test_second_fill_on_the_same_objects(oracle):
    """Forces::fill called again on the same Mesh / Forces after a position update (what Cloth::step does every step)."""
    X, fn, x = _case("regular2_12")
    x2 = x + 1e-3 * np.random.default_rng(3).standard_normal(x.shape)
    a, b = oracle.ref_forces_fill(fn, x, X, more_steps=[(x2, None)])
    es = oracle.arcsim_edge_stencils(X.shape[0], fn)
    _compare(a, oracle.forces_fill(fn, es, x, X), X.shape[0], "first fill")
    _compare(b, oracle.forces_fill(fn, es, x2, X), X.shape[0], "second fill")
    assert a["M"][2].tobytes() == b["M"][2].tobytes() and a["MDK"][2].tobytes() != b["MDK"][2].tobytes()


def test_pattern_sizes_counted_by_the_reference(oracle):
    """regular2 n=3: nnz(M)=369, nnz(MDK)=513; build4 n=3: 621 / 837 — and the explicit zeros of the mass blocks
    (ComputeInertial.cpp:45-46) survive setFromTriplets: two thirds of M's stored entries are zeros."""
    for gen, nM, nK in (("regular2", 369, 513), ("build4", 621, 837)):
        X, fn = getattr(E.meshgen, gen)(3)
        r = oracle.ref_forces_fill(fn, E.meshgen.drape_state(X, seed=0), X)
        assert (len(r["M"][2]), len(r["MDK"][2])) == (nM, nK)
        assert np.count_nonzero(r["M"][2] == 0.0) == 2 * nM // 3
        for k in ("M", "MDK"):        # sorted inner indices, symmetric pattern and (exactly) symmetric values: mirrored triplets
            outer, inner, vals = r[k]
            import scipy.sparse as sp
            A = sp.csc_matrix((vals, inner, outer), shape=(r["dof"], r["dof"]))
            assert all(np.all(np.diff(inner[outer[c]:outer[c + 1]]) > 0) for c in range(r["dof"]))
            assert abs(A - A.T).max() == 0.0


@pytest.mark.parametrize("name", ["regular2_12", "build4_7", "delaunay_300"])
def test_normals_of_compute_ws_data(oracle, name):
    """face->n / node->n as ArcSim's own compute_ws_data leaves them (mesh.cpp:135-151, geometry.cpp:302-316) == oracle.mesh_normals,
    bit for bit — also after positions moved (Cloth::step: x += h v, then compute_ws_data, Cloth.cpp:394-410)."""
    X, fn, x = _case(name)
    es, fnrm, nnrm = oracle.ref_mesh_data(fn, x, X)
    f0, n0 = oracle.mesh_normals(fn, x)
    assert fnrm.tobytes() == f0.tobytes() and nnrm.tobytes() == n0.tobytes()
    x2 = x + 2e-3 * np.random.default_rng(5).standard_normal(x.shape)
    _, fnrm2, nnrm2 = oracle.ref_mesh_data(fn, x, X, x_new=x2)
    f2, n2 = oracle.mesh_normals(fn, x2)
    assert fnrm2.tobytes() == f2.tobytes() and nnrm2.tobytes() == n2.tobytes()


def test_golden_fixtures_come_from_the_reference(oracle):
    """tests/golden/forces_*.npz are written from libforces_ref.so (make_golden.py); the restatement reproduces them."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for fname, gen, n, seed in (("forces_regular2_n12.npz", "regular2", 12, 0), ("forces_build4_n7.npz", "build4", 7, 1)):
        g = np.load(os.path.join(gold, fname))
        assert str(g["source"]) == "libforces_ref"
        X, fn = getattr(E.meshgen, gen)(n)
        x = E.meshgen.drape_state(X, seed=seed)
        r = oracle.ref_forces_fill(fn, x, X)
        assert r["MDK"][2].tobytes() == g["K_vals"].tobytes() and r["M"][2].tobytes() == g["M_vals"].tobytes() and r["f"].tobytes() == g["f"].tobytes()
