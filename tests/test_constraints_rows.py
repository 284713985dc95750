"""Contact part of Constraints::fill (Constraints.cpp:424-468): the inequality rows built from a CD2 contact list.
Integer columns and FP64 values (one negation / one multiply each) are compared bit for bit with the restatement in the oracle."""
import ctypes
import os

import numpy as np
import pytest

import eol_cloth_b200 as E
from eol_cloth_b200 import capi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _compare(rows_ref, nnz, cols, vals):
    assert len(rows_ref) == len(nnz)
    for r, ref in enumerate(rows_ref):
        assert nnz[r] == len(ref)
        assert [int(c) for c in cols[r, :nnz[r]]] == [c for c, _ in ref]
        assert np.array(vals[r, :nnz[r]]).tobytes() == np.array([v for _, v in ref], dtype=np.float64).tobytes()
        assert np.all(cols[r, nnz[r]:] == -1) and np.all(vals[r, nnz[r]:] == 0.0)


@pytest.mark.parametrize("name", ["cd_build4_n16_corner", "cd_regular2_n24"])
def test_host_rows_match_oracle_on_golden_contacts(oracle, name):
    """Host entry point (pure host code of the library: no device needed) on the golden CD2 lists: all three contact kinds."""
    c = np.load(os.path.join(GOLD, name + ".npz"))["cd2"]
    c = np.ascontiguousarray(c.view(E.CONTACT_DTYPE) if c.dtype != E.CONTACT_DTYPE else c).reshape(-1)
    kinds = {(int(a), int(b)) for a, b in zip(c["count1"], c["count2"])}
    if name.endswith("corner"):
        assert kinds == {(3, 1), (1, 3), (2, 2)}
    nnz, cols, vals = E.contact_rows(c)
    _compare(oracle.constraints_contact_rows(c), nnz, cols, vals)
    # EoL nodes: contacts touching one take no row and the numbering closes up
    N = int(max(c["verts2"].max(), 0)) + 1
    eol = np.zeros(N, np.uint8)
    eol[c["verts2"][::3, 0]] = 1
    nnz2, cols2, vals2 = E.contact_rows(c, eol)
    ref2 = oracle.constraints_contact_rows(c, eol.astype(bool))
    assert 0 < len(ref2) < len(nnz)
    _compare(ref2, nnz2, cols2, vals2)
    assert E.contact_rows(c[:0])[0].size == 0


def test_fixed_corner_rows(oracle):
    """eolc_constraints_fixed_rows == the fixed-corner part of Constraints::fill (Constraints.cpp:470-497), host only."""
    import ctypes
    from eol_cloth_b200 import capi
    L = capi.lib()
    rng = np.random.default_rng(3)
    v = rng.standard_normal((50, 3))
    cases = [
        (np.array([[1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0], [-1, 0, 0, 0, 0, 0], [-1, 0, 0, 0, 0, 0]], float), [3, 4, 0, 0], 0),   # simulationSettings.json: corners 3, 4 pinned
        (np.array([[1, 0, 1, 0.1, 0.2, 0.3], [0, 0, 0, 1, 1, 1], [-1, 1, 1, 9, 9, 9], [0, 1, 0, -0.5, 0.25, 2]], float), [0, 7, 49, 21], 5),
        (np.full((4, 6), -1.0), [0, 0, 0, 0], 2),
    ]
    for c, ci, row0 in cases:
        ci = np.array(ci, np.int32)
        n = ctypes.c_int32(0)
        rows, cols, vals, beq = np.zeros(12, np.int32), np.zeros(12, np.int32), np.zeros(12), np.zeros(12)
        capi.check(L.eolc_constraints_fixed_rows(capi.dptr(np.ascontiguousarray(c)), capi.iptr(ci), capi.dptr(np.ascontiguousarray(v)), 50, row0,
                                                 ctypes.byref(n), capi.iptr(rows), capi.iptr(cols), capi.dptr(vals), capi.dptr(beq)))
        r, cc, vv, bb = oracle.constraints_fixed_rows(c, ci, v, row0)
        assert n.value == len(r)
        assert np.array_equal(rows[:n.value], r) and np.array_equal(cols[:n.value], cc)
        assert vals[:n.value].tobytes() == vv.tobytes() and beq[:n.value].tobytes() == bb.tobytes()
    assert L.eolc_constraints_fixed_rows(capi.dptr(np.ascontiguousarray(cases[0][0])), capi.iptr(np.array([3, 99, 0, 0], np.int32)),
                                         capi.dptr(np.ascontiguousarray(v)), 50, 0, ctypes.byref(n), capi.iptr(rows), capi.iptr(cols), capi.dptr(vals), capi.dptr(beq)) == -1


def _compare_csr(rows_ref, row_ptr, cols, vals):
    """CSR rows == the oracle's rows, entry for entry (the reference's triplet sequence), bit for bit."""
    assert len(row_ptr) == len(rows_ref) + 1 and row_ptr[0] == 0
    assert np.array_equal(np.diff(row_ptr), [len(r) for r in rows_ref])
    assert [int(c) for c in cols] == [c for r in rows_ref for c, _ in r]
    assert np.asarray(vals).tobytes() == np.array([v for r in rows_ref for _, v in r], dtype=np.float64).tobytes()


@pytest.mark.gpu
def test_device_rows_of_last_run(ctx, oracle):
    from eol_cloth_b200.collisions import make_obstacles
    X, fn = E.meshgen.regular2(40)
    c0 = np.array([0.9175, -0.25, -0.549])
    x = E.meshgen.box_scene_state(X, seed=3, centre=c0)
    obs = make_obstacles(E.meshgen.BOX_THRESHOLD, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c0)[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, E.meshgen.BOX_THRESHOLD)
    contacts = plan.run(x, obs, 0, 0)
    assert len(contacts) > 50
    ref = oracle.constraints_contact_rows(contacts)
    _compare(ref, *plan.contact_rows())
    _compare(ref, *E.contact_rows(contacts))
    _compare_csr(ref, *plan.contact_rows_csr())
    eol = np.zeros(X.shape[0], np.uint8)
    eol[contacts["verts2"][::2, 0]] = 1
    ref_eol = oracle.constraints_contact_rows(contacts, eol.astype(bool))
    assert len(ref_eol) < len(ref)
    _compare(ref_eol, *plan.contact_rows(eol))
    _compare_csr(ref_eol, *plan.contact_rows_csr(eol))
    # pageable result arrays and the capacity error of the C entry
    L, n, nz = capi.lib(), ctypes.c_int32(0), ctypes.c_int32(0)
    rp, cc, vv = np.zeros(len(contacts) + 1, np.int32), np.zeros(9 * len(contacts), np.int32), np.zeros(9 * len(contacts))
    assert L.eolc_cd_contact_rows_csr(plan._h, None, len(contacts), 9 * len(contacts), ctypes.byref(n), ctypes.byref(nz), capi.iptr(rp),
                                      capi.iptr(cc), capi.dptr(vv)) == 0
    _compare_csr(ref, rp[:n.value + 1], cc[:nz.value], vv[:nz.value])
    assert L.eolc_cd_contact_rows_csr(plan._h, None, 3, 9, ctypes.byref(n), ctypes.byref(nz), capi.iptr(rp), capi.iptr(cc), capi.dptr(vv)) == -3
    assert n.value == len(contacts) and nz.value == sum(len(r) for r in ref)


@pytest.mark.gpu
def test_resident_run_feeds_device_rows(ctx, oracle):
    """eolc_cd_run_batched_resident_dev: the records never leave the device; offsets and the rows built from them on the device equal
    those of the host-copied list (3 scenes)."""
    import torch
    from eol_cloth_b200.collisions import make_obstacles
    X, fn = E.meshgen.regular2(36)
    c0 = np.array([0.9175, -0.25, -0.549])
    obs = make_obstacles(E.meshgen.BOX_THRESHOLD, box_whd=E.meshgen.BOX_WHD[None], box_E=E.meshgen.box_frame(c0)[None])
    plan = E.CollisionPlan(ctx, X.shape[0], fn, E.meshgen.BOX_THRESHOLD)
    xs = np.stack([E.meshgen.box_scene_state(X, seed=s, centre=c0) for s in range(3)])
    xd = torch.from_numpy(xs).to(torch.device("cuda", ctx.device))
    torch.cuda.synchronize()
    host_list, off_h = plan.run(xd.data_ptr(), obs, 0, 0, x_is_device_ptr=True, n_scenes=3)
    off_r = plan.run_resident(xd.data_ptr(), obs, 0, 0, n_scenes=3)
    assert np.array_equal(off_r, off_h)
    ptr, n = plan.contacts_dev()
    assert n == len(host_list) and ptr
    _compare(oracle.constraints_contact_rows(host_list), *plan.contact_rows())
    _compare_csr(oracle.constraints_contact_rows(host_list), *plan.contact_rows_csr())


# ---- against THE REFERENCE'S OWN Constraints::fill (oracle/_ref/libconstraints_ref.so: Constraints.cpp, Collisions.cpp, boxTriCollision.cpp,
# Box.cpp, Obstacles.cpp, ... compiled unmodified; it calls CD2 itself and assembles Aineq / Aeq with setFromTriplets) ---------------
def _csc_of_rows(n_rows, n_cols, nnz, cols, vals, row0=0):
    """The library's rows as the compressed column-major matrix Eigen builds from the same triplets."""
    import scipy.sparse as sp
    r = np.repeat(np.arange(len(nnz)), nnz) + row0
    keep = np.arange(9)[None, :] < np.asarray(nnz)[:, None]
    A = sp.csc_matrix((np.asarray(vals).reshape(-1, 9)[keep], (r, np.asarray(cols).reshape(-1, 9)[keep])), shape=(n_rows, n_cols))
    A.sort_indices()
    return A


def _assert_same_csc(A, ref, what):
    rows, outer, inner, vals = ref
    assert A.shape[0] == rows, what
    assert np.array_equal(A.indptr, outer) and np.array_equal(A.indices, inner), what + ": index arrays"
    assert A.data.tobytes() == vals.tobytes(), what + ": values differ from the reference's"


@pytest.mark.parametrize("gen,n,centre,points", [("regular2", 24, None, False), ("build4", 16, (0.9175, -0.25, -0.549), True), ("regular2", 40, (0.9175, -0.25, -0.549), False)])
def test_rows_equal_the_reference_constraints_fill(oracle, gen, n, centre, points):
    X, fn = getattr(E.meshgen, gen)(n)
    N = X.shape[0]
    centre = E.meshgen.BOX_CENTRE if centre is None else np.asarray(centre)
    x = E.meshgen.box_scene_state(X, seed=n, centre=centre)
    v = 0.1 * np.random.default_rng(n).standard_normal((N, 3))
    pxyz = pn = None
    if points:
        pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3])
        pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]])
    whd, Em = E.meshgen.BOX_WHD[None], E.meshgen.box_frame(centre)[None]
    fixed_c = np.array([[1, 1, 1, 0.0, 0.0, 0.1], [1, 0, 1, 0.2, 0.0, 0.0], [-1, 0, 0, 0, 0, 0], [0, 1, 0, 0, -0.3, 0]], float)
    fixed_ci = np.array([0, n - 1, 3, N - 1], np.int32)
    ref = oracle.ref_constraints_fill(fn, x, X, v, E.meshgen.BOX_THRESHOLD, pxyz, pn, whd, Em, fixed_c, fixed_ci)
    contacts = oracle.ref_cd(fn, x, E.meshgen.BOX_THRESHOLD, pxyz, pn, whd, Em, which=0)     # the CD2 list the reference built inside
    assert ref["hasCollisions"] and ref["hasFixed"] and len(contacts) > 20
    nnz, cols, vals = E.contact_rows(contacts)
    assert len(nnz) == ref["Aineq"][0] == len(contacts)
    _assert_same_csc(_csc_of_rows(len(nnz), 3 * N, nnz, cols, vals), ref["Aineq"], "Aineq")
    assert not ref["bineq"].any()
    # the oracle's restatement too
    rr = oracle.constraints_contact_rows(contacts)
    assert [len(r) for r in rr] == list(nnz)
    # fixed corners -> Aeq / beq
    L, m = capi.lib(), ctypes.c_int32(0)
    rows, cc, vv, bb = np.zeros(12, np.int32), np.zeros(12, np.int32), np.zeros(12), np.zeros(12)
    capi.check(L.eolc_constraints_fixed_rows(capi.dptr(np.ascontiguousarray(fixed_c)), capi.iptr(fixed_ci), capi.dptr(np.ascontiguousarray(v)), N, 0,
                                             ctypes.byref(m), capi.iptr(rows), capi.iptr(cc), capi.dptr(vv), capi.dptr(bb)))
    import scipy.sparse as sp
    A = sp.csc_matrix((vv[:m.value], (rows[:m.value], cc[:m.value])), shape=(m.value, 3 * N))
    A.sort_indices()
    _assert_same_csc(A, ref["Aeq"], "Aeq")
    assert bb[:m.value].tobytes() == ref["beq"].tobytes()


def test_rows_with_eol_nodes_equal_the_reference(oracle):
    """Contacts touching an EoL node take no row (`continue`, :426,:440,:454).  The EoL nodes are given Node::cornerID = an obstacle
    point, for which the reference first adds its own rows (one inequality, two equalities per node, :144-211): the contact rows
    then start at row #EoL and the fixed rows at 2 #EoL."""
    X, fn = E.meshgen.regular2(30)
    N = X.shape[0]
    c0 = np.array([0.9175, -0.25, -0.549])
    x = E.meshgen.box_scene_state(X, seed=5, centre=c0)
    v = np.zeros((N, 3))
    pxyz, pn = np.array([[0.1, 0.8, -0.2]]), np.array([[0.0, 0.6, 0.8]])
    whd, Em = E.meshgen.BOX_WHD[None], E.meshgen.box_frame(c0)[None]
    contacts = oracle.ref_cd(fn, x, E.meshgen.BOX_THRESHOLD, pxyz, pn, whd, Em, which=0)
    eol_nodes = np.unique(contacts["verts2"][::3, 0])[:25]
    corner_id = np.full(N, -1, np.int32); corner_id[eol_nodes] = 0
    fixed_c = np.array([[1, 1, 1, 0, 0, 0], [-1, 0, 0, 0, 0, 0], [-1, 0, 0, 0, 0, 0], [-1, 0, 0, 0, 0, 0.0]])
    ref = oracle.ref_constraints_fill(fn, x, X, v, E.meshgen.BOX_THRESHOLD, pxyz, pn, whd, Em, fixed_c, np.array([1, 0, 0, 0], np.int32), corner_id=corner_id)
    flags = np.zeros(N, np.uint8); flags[eol_nodes] = 1
    nnz, cols, vals = E.contact_rows(contacts, flags)
    ne = len(eol_nodes)
    assert 0 < len(nnz) < len(contacts) and ref["Aineq"][0] == ne + len(nnz) and ref["Aeq"][0] == 2 * ne + 3
    import scipy.sparse as sp
    rows, outer, inner, valsr = ref["Aineq"]
    R = sp.csc_matrix((valsr, inner, outer), shape=(rows, 3 * N + 2 * ne)).tocsr()[ne:, :3 * N].tocsc()
    R.sort_indices()
    A = _csc_of_rows(len(nnz), 3 * N, nnz, cols, vals)
    assert np.array_equal(A.indptr, R.indptr) and np.array_equal(A.indices, R.indices) and A.data.tobytes() == R.data.tobytes()
