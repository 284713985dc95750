"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (eol_cloth_b200) never does.
"""
import ctypes
import os
import time
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)


class Contact(ctypes.Structure):
    """POD mirror of include/eolc.h eolc_contact (btc::Collision, boxTriCollision.h:49-110)."""
    _fields_ = [
        ("dist", ctypes.c_double),
        ("nor1", ctypes.c_double * 3), ("nor2", ctypes.c_double * 3),
        ("pos1", ctypes.c_double * 3), ("pos2", ctypes.c_double * 3), ("pos1_", ctypes.c_double * 3),
        ("weights1", ctypes.c_double * 3), ("weights2", ctypes.c_double * 3),
        ("edgeDir", ctypes.c_double * 3),
        ("count1", ctypes.c_int32), ("count2", ctypes.c_int32),
        ("verts1", ctypes.c_int32 * 3), ("verts2", ctypes.c_int32 * 3),
        ("tri1", ctypes.c_int32), ("tri2", ctypes.c_int32),
        ("edge1", ctypes.c_int32 * 3),
        ("n_edge1", ctypes.c_int32),
        ("edge2", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


CONTACT_DTYPE = np.dtype([
    ("dist", "f8"), ("nor1", "f8", 3), ("nor2", "f8", 3), ("pos1", "f8", 3), ("pos2", "f8", 3), ("pos1_", "f8", 3),
    ("weights1", "f8", 3), ("weights2", "f8", 3), ("edgeDir", "f8", 3),
    ("count1", "i4"), ("count2", "i4"), ("verts1", "i4", 3), ("verts2", "i4", 3), ("tri1", "i4"), ("tri2", "i4"),
    ("edge1", "i4", 3), ("n_edge1", "i4"), ("edge2", "i4"), ("reserved", "i4"),
])
assert CONTACT_DTYPE.itemsize == ctypes.sizeof(Contact) == 264


def build():
    """Compile oracle/liboracle.so (and oracle/_ref/ when /root/reference is present)."""
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = ctypes.CDLL(path)
    L.oracle_forces_fill.restype = ctypes.c_void_p
    L.oracle_forces_fill.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, ctypes.c_int, c_ip, c_dp, c_dp, c_dp, c_dp,
                                     ctypes.c_double, ctypes.c_int]
    L.oracle_forces_fill_eol.restype = ctypes.c_void_p
    L.oracle_forces_fill_eol.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, ctypes.c_int, c_ip, c_dp, c_dp, c_dp, c_dp,
                                         ctypes.c_double, c_ip, ctypes.c_int]
    L.oracle_forces_dof.argtypes = [ctypes.c_void_p]
    L.oracle_forces_f.restype = c_dp
    L.oracle_forces_f.argtypes = [ctypes.c_void_p]
    L.oracle_forces_nnz.restype = ctypes.c_int64
    L.oracle_forces_nnz.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.oracle_forces_outer.restype = c_ip
    L.oracle_forces_outer.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.oracle_forces_inner.restype = c_ip
    L.oracle_forces_inner.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.oracle_forces_vals.restype = c_dp
    L.oracle_forces_vals.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.oracle_forces_seconds.restype = ctypes.c_double
    L.oracle_forces_seconds.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.oracle_forces_free.argtypes = [ctypes.c_void_p]
    L.oracle_cd.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, ctypes.c_double, ctypes.c_int, c_dp, c_dp,
                            ctypes.c_int, c_dp, c_dp, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                            ctypes.POINTER(ctypes.c_int)]
    L.oracle_cd_edges.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, c_ip, c_dp]
    L.oracle_perturbation.argtypes = [ctypes.c_int, ctypes.c_double, c_dp]
    L.oracle_rng_raw.argtypes = [ctypes.c_int, c_dp]
    L.oracle_raytri.argtypes = [c_dp] * 6
    L.oracle_box.argtypes = [c_dp] * 6
    L.oracle_face_frame.argtypes = [c_dp] * 8
    _LIB = L
    return L


def _d(a):
    return a.ctypes.data_as(c_dp)


def _i(a):
    return a.ctypes.data_as(c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---- synthetic inputs without the product package (bench.py --impl reference must not load libeolc_b200.so) -------------------
def arcsim_edge_stencils(n_nodes, face_nodes):
    """(E,4) int32 stencils (n0, n1, opp(adjf[0]), opp(adjf[1])), -1 = no face, in the order ArcSim creates mesh.edges while faces
    are added (Mesh::add(Face), src/external/ArcSim/mesh.cpp:356-378): for face k, i = 0,1,2, edge (v[i] -> v[i+1]) is appended iff
    no edge joins those nodes yet; adjf[side] with side = 0 iff the face runs n[0] -> n[1].  Vectorised numpy restatement;
    tests/test_oracle.py holds it equal to the loop restatement and to eolc_mesh_edge_stencils."""
    fn = _i32(face_nodes).reshape(-1, 3)
    a = fn[:, [0, 1, 2]].ravel()
    b = fn[:, [1, 2, 0]].ravel()
    c = fn[:, [2, 0, 1]].ravel()
    key = np.minimum(a, b).astype(np.int64) * int(n_nodes) + np.maximum(a, b)
    order = np.argsort(key, kind="stable")
    ks = key[order]
    first = np.ones(ks.size, bool)
    first[1:] = ks[1:] != ks[:-1]
    gid = np.cumsum(first) - 1
    creators = order[first]                  # the half-edge that creates each edge = its first occurrence
    eorder = np.argsort(creators, kind="stable")
    rank = np.empty(eorder.size, np.int64)
    rank[eorder] = np.arange(eorder.size)
    ce = creators[eorder]
    es = np.full((ce.size, 4), -1, np.int32)
    es[:, 0], es[:, 1], es[:, 2] = a[ce], b[ce], c[ce]
    later = order[~first]
    e2 = rank[gid[~first]]
    side = np.where(a[later] == es[e2, 0], 0, 1)
    es[e2, 2 + side] = c[later]
    return es


def sheet_regular2(n, m=None, rows=None):
    """regular2 sheet of the benchmarks: node (i, j) -> i m + j at X = (i/(n-1), j/(m-1)), two counter-clockwise
    triangles per cell.  rows: only the first `rows` grid rows (a strip of the same sheet: same coordinates, same numbering)."""
    m = n if m is None else m
    r = n if rows is None else rows
    i, j = np.meshgrid(np.arange(r), np.arange(m), indexing="ij")
    X = np.stack([i.ravel() / (n - 1), j.ravel() / (m - 1)], axis=1).astype(np.float64)
    ci, cj = np.meshgrid(np.arange(r - 1), np.arange(m - 1), indexing="ij")
    k0 = (ci * m + cj).ravel()
    faces = np.empty((2 * k0.size, 3), dtype=np.int32)
    faces[0::2] = np.stack([k0, k0 + m, k0 + m + 1], axis=1)
    faces[1::2] = np.stack([k0, k0 + m + 1, k0 + 1], axis=1)
    return X, faces


def drape_state(X, seed=0, amp=0.05, noise=1e-3, n_total=None):
    """x = (X, 0.05 sin(2 pi X0) cos(2 pi X1)) + U(-1e-3, 1e-3) on all coords; n_total: the random stream
    is drawn for that many nodes and the first len(X) are used (a strip gets the states of the full sheet's nodes)."""
    rng = np.random.default_rng(seed)
    r = rng.uniform(-noise, noise, size=((X.shape[0] if n_total is None else n_total), 3))[:X.shape[0]]
    x = np.zeros((X.shape[0], 3))
    x[:, :2] = X
    x[:, 2] = amp * np.sin(2 * np.pi * X[:, 0]) * np.cos(2 * np.pi * X[:, 1])
    return x + r


MATERIAL_DEFAULT = (0.05, 50.0, 0.01, 1.0e-5, 0.0, 1.0)  # simulationSettings.json:27-33
GRAV_DEFAULT = (0.0, 0.0, -9.8)
H_DEFAULT = 0.5e-2


def forces_fill(face_nodes, edge_stencil, x, X, mat=MATERIAL_DEFAULT, grav=GRAV_DEFAULT, h=H_DEFAULT,
                skip_assembly=False, eol_index=None, threads=1):
    """Reference Forces::fill on flat arrays.  Returns dict(f, M=(outer, inner, vals), MDK=..., seconds=(el, asm)).
    eol_index (N ints, -1 = Lagrangian, k >= 0 = Node::EoL_index) switches the touched elements to the EOL branch.
    threads > 1: TIMING variant for bench.py's CPU legs (the reference is single-threaded): the element loops run on that many
    threads and the two setFromTriplets side by side; the triplet order and the sums are those of one thread, bit for bit."""
    L = lib()
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    edge_stencil = _i32(edge_stencil).reshape(-1, 4)
    x = _f64(x).reshape(-1, 3)
    X = _f64(X).reshape(-1, 2)
    N = x.shape[0]
    matv = _f64(mat)
    gv = _f64(grav)
    eol = None if eol_index is None else _i32(eol_index).reshape(N)
    r = L.oracle_forces_fill_eol(N, face_nodes.shape[0], _i(face_nodes), edge_stencil.shape[0], _i(edge_stencil), _d(x),
                                 _d(X), _d(matv), _d(gv), float(h), None if eol is None else _i(eol),
                                 (1 if skip_assembly else 0) | (max(1, min(int(threads), 255)) << 8))
    if not r:
        raise RuntimeError("oracle_forces_fill_eol: a bending stencil names a face the mesh does not have")
    try:
        dof = L.oracle_forces_dof(r)
        out = {"dof": dof, "f": np.ctypeslib.as_array(L.oracle_forces_f(r), (dof,)).copy(),
               "seconds": (L.oracle_forces_seconds(r, 0), L.oracle_forces_seconds(r, 1))}
        if not skip_assembly:
            for which, name in ((0, "M"), (1, "MDK")):
                nnz = L.oracle_forces_nnz(r, which)
                outer = np.ctypeslib.as_array(L.oracle_forces_outer(r, which), (dof + 1,)).copy()
                inner = np.ctypeslib.as_array(L.oracle_forces_inner(r, which), (nnz,)).copy()
                vals = np.ctypeslib.as_array(L.oracle_forces_vals(r, which), (nnz,)).copy()
                out[name] = (outer, inner, vals)
    finally:
        L.oracle_forces_free(r)
    return out


def cd(face_nodes, x, threshold, pxyz=None, pnorms=None, box_whd=None, box_E=None, point_eol_flag=0, remap=0,
       capacity=None):
    """Reference CD (point_eol_flag=1, remap=1) / CD2 (0, 0) on flat arrays -> structured array of contacts."""
    L = lib()
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    N, F = x.shape[0], face_nodes.shape[0]
    pxyz = _f64(np.zeros((0, 3)) if pxyz is None else pxyz).reshape(-1, 3)
    pnorms = _f64(np.zeros((0, 3)) if pnorms is None else pnorms).reshape(-1, 3)
    box_whd = _f64(np.zeros((0, 3)) if box_whd is None else box_whd).reshape(-1, 3)
    box_E = _f64(np.zeros((0, 16)) if box_E is None else box_E).reshape(-1, 16)
    if capacity is None:
        capacity = N + 8 * (1 + box_whd.shape[0]) + pxyz.shape[0] + 4096
    while True:
        out = np.zeros(capacity, dtype=CONTACT_DTYPE)
        n = ctypes.c_int(0)
        rc = L.oracle_cd(N, F, _i(face_nodes), _d(x), float(threshold), pxyz.shape[0], _d(pxyz), _d(pnorms),
                         box_whd.shape[0], _d(box_whd), _d(box_E), int(point_eol_flag), int(remap),
                         out.ctypes.data_as(ctypes.c_void_p), capacity, ctypes.byref(n))
        if rc == 0:
            return out[:n.value].copy()
        capacity = n.value


# ---- the reference's own collision code (oracle/_ref/libbtc_ref*.so: boxTriCollision.cpp, Collisions.cpp, raytri.cpp compiled
# UNMODIFIED against oracle/mini_eigen by oracle/Makefile; prebuilt files travel to the GPU box) ---------------------------------
_REFLIBS = {}


def ref_lib(scalar_redux=False, adapter=False):
    """libbtc_ref.so (Eigen 3.3 SSE2 reduction order), libbtc_ref_scalar.so (non-vectorised order), or — adapter=True —
    libadapter_cd.so: the same driver and reference types, but CD / CD2 are this repository's adapter/Collisions_b200.cpp, i.e. the
    call lands on the GPU through include/eolc_host.hpp and the C ABI (needs a CUDA device; aborts like the reference otherwise)."""
    name = "libadapter_cd.so" if adapter else "libbtc_ref_scalar.so" if scalar_redux else "libbtc_ref.so"
    if name in _REFLIBS:
        return _REFLIBS[name]
    path = os.path.join(_HERE, "_ref", name)
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and /root/reference is not here to build it from")
    L = ctypes.CDLL(path)
    L.ref_cd.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, c_ip, ctypes.c_double, ctypes.c_int, c_dp, c_dp, ctypes.c_int,
                         c_dp, c_dp, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.ref_btc_box.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, ctypes.c_double, c_dp, c_dp, ctypes.c_int,
                              ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.ref_btc_points.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, ctypes.c_double, ctypes.c_int, c_dp, c_dp, ctypes.c_int,
                                 ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    L.ref_btc_edges.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, c_ip, c_dp, c_ip, c_dp]
    L.ref_btc_boxtables.argtypes = [c_dp] * 6
    L.ref_btc_hash_defined.argtypes = [ctypes.c_int, ctypes.c_int]
    assert adapter or L.ref_redux_order() == (1 if scalar_redux else 0)
    _REFLIBS[name] = L
    return L


def ref_hash_defined(N, F):
    """True while the reference's `int` edge hash (boxTriCollision.cpp:167-169) does not overflow: (3F+1)(N) + N < 2^31."""
    return (3 * F + 1) * N + N < 2 ** 31 - 1


def _ref_call(fn, N, *args):
    capacity = N + 4096
    while True:
        out = np.zeros(capacity, dtype=CONTACT_DTYPE)
        n = ctypes.c_int(0)
        rc = fn(*args, out.ctypes.data_as(ctypes.c_void_p), capacity, ctypes.byref(n))
        if rc == 0:
            return out[:n.value].copy()
        capacity = n.value


def ref_cd(face_nodes, x, threshold, pxyz=None, pnorms=None, box_whd=None, box_E=None, which=0, eol=None, scalar_redux=False,
           adapter=False):
    """The reference's CD (which=1, Collisions.cpp:11-53) / CD2 (which=0, :55-78), run as compiled from its own sources; with
    adapter=True the same call sites reach adapter/Collisions_b200.cpp instead (the drop-in, executed)."""
    L = ref_lib(scalar_redux, adapter)
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    N, F = x.shape[0], face_nodes.shape[0]
    if not adapter and box_whd is not None and len(box_whd) and not ref_hash_defined(N, F):
        raise ValueError("the reference's int edge hash overflows (undefined behaviour) for this mesh size")
    pxyz = _f64(np.zeros((0, 3)) if pxyz is None else pxyz).reshape(-1, 3)
    pnorms = _f64(np.zeros((0, 3)) if pnorms is None else pnorms).reshape(-1, 3)
    box_whd = _f64(np.zeros((0, 3)) if box_whd is None else box_whd).reshape(-1, 3)
    box_E = _f64(np.zeros((0, 16)) if box_E is None else box_E).reshape(-1, 16)
    eolv = None if eol is None else _i32(eol).reshape(N)
    return _ref_call(L.ref_cd, N + 8 * box_whd.shape[0] + pxyz.shape[0], N, F, _i(face_nodes), _d(x),
                     None if eolv is None else _i(eolv), float(threshold), pxyz.shape[0], _d(pxyz), _d(pnorms),
                     box_whd.shape[0], _d(box_whd), _d(box_E), int(which))


def ref_btc_box(face_nodes, x, threshold, whd, E1, EOL=False, scalar_redux=False):
    """btc::boxTriCollision (boxTriCollision.cpp:603-1063) for one box; EOL=True skips section (A) (:674)."""
    L = ref_lib(scalar_redux)
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    whd = _f64(whd).reshape(3)
    E1 = _f64(E1).reshape(16)
    return _ref_call(L.ref_btc_box, x.shape[0], x.shape[0], face_nodes.shape[0], _i(face_nodes), _d(x), float(threshold),
                     _d(whd), _d(E1), int(bool(EOL)))


def ref_btc_points(face_nodes, x, threshold, pxyz, pnorms, EOL=False, scalar_redux=False):
    """btc::pointTriCollision (boxTriCollision.cpp:1067-1224)."""
    L = ref_lib(scalar_redux)
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    pxyz = _f64(pxyz).reshape(-1, 3)
    pnorms = _f64(pnorms).reshape(-1, 3)
    return _ref_call(L.ref_btc_points, x.shape[0] + pxyz.shape[0], x.shape[0], face_nodes.shape[0], _i(face_nodes), _d(x),
                     float(threshold), pxyz.shape[0], _d(pxyz), _d(pnorms), int(bool(EOL)))


def ref_btc_edges(face_nodes, x, scalar_redux=False):
    """btc::createEdges (boxTriCollision.cpp:141-231) -> (table [E,6] = verts(4)+faces(2), normals [E,6], internal [E], angle [E])."""
    L = ref_lib(scalar_redux)
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    F = face_nodes.shape[0]
    if not ref_hash_defined(x.shape[0], F):
        raise ValueError("the reference's int edge hash overflows (undefined behaviour) for this mesh size")
    tab = np.zeros((3 * F, 6), dtype=np.int32)
    nrm = np.zeros((3 * F, 6), dtype=np.float64)
    internal = np.zeros(3 * F, dtype=np.int32)
    angle = np.zeros(3 * F, dtype=np.float64)
    E = L.ref_btc_edges(x.shape[0], F, _i(face_nodes), _d(x), _i(tab), _d(nrm), _i(internal), _d(angle))
    return tab[:E].copy(), nrm[:E].copy(), internal[:E].copy(), angle[:E].copy()


def ref_btc_boxtables(whd, E1, scalar_redux=False):
    """btc::createBox (boxTriCollision.cpp:400-420): verts1 (14,3), faceNors1 (24,3), vertNors1 (14,3), box edge angles (12)."""
    L = ref_lib(scalar_redux)
    whd = _f64(whd).reshape(3)
    E1 = _f64(E1).reshape(16)
    v, fnr, vn, ang = np.zeros((14, 3)), np.zeros((24, 3)), np.zeros((14, 3)), np.zeros(12)
    L.ref_btc_boxtables(_d(whd), _d(E1), _d(v), _d(fnr), _d(vn), _d(ang))
    return v, fnr, vn, ang


def cd_edges(face_nodes, x):
    L = lib()
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    F = face_nodes.shape[0]
    tab = np.zeros((3 * F, 6), dtype=np.int32)
    nrm = np.zeros((3 * F, 6), dtype=np.float64)
    E = L.oracle_cd_edges(x.shape[0], F, _i(face_nodes), _d(x), _i(tab), _d(nrm))
    return tab[:E].copy(), nrm[:E].copy()


def perturbation(n, threshold):
    out = np.zeros((n, 3))
    lib().oracle_perturbation(n, float(threshold), _d(out))
    return out


def rng_raw(n):
    out = np.zeros(n)
    lib().oracle_rng_raw(n, _d(out))
    return out


def raytri(orig, dirv, v0, v1, v2):
    tuv = np.zeros(3)
    hit = lib().oracle_raytri(_d(_f64(orig)), _d(_f64(dirv)), _d(_f64(v0)), _d(_f64(v1)), _d(_f64(v2)), _d(tuv))
    return hit, tuv


def box(whd, E1):
    v, fn, vn, ang = np.zeros((14, 3)), np.zeros((24, 3)), np.zeros((14, 3)), np.zeros(12)
    lib().oracle_box(_d(_f64(whd)), _d(_f64(E1)), _d(v), _d(fn), _d(vn), _d(ang))
    return v, fn, vn, ang


def compute_membrane(xa, xb, xc, Xa, Xb, Xc, e, nu, P, Q):
    L = lib()
    L.oracle_compute_membrane.argtypes = [c_dp] * 6 + [ctypes.c_double] * 2 + [c_dp] * 5
    W, f, K = np.zeros(1), np.zeros(9), np.zeros(81)
    L.oracle_compute_membrane(*[_d(_f64(a)) for a in (xa, xb, xc, Xa, Xb, Xc)], e, nu, _d(_f64(P)), _d(_f64(Q)),
                              _d(W), _d(f), _d(K))
    return W[0], f, K.reshape(9, 9)


def compute_bending(x0, x1, x2, x3, X0, X1, X2, X3, beta):
    L = lib()
    L.oracle_compute_bending.argtypes = [c_dp] * 8 + [ctypes.c_double] + [c_dp] * 3
    W, f, K = np.zeros(1), np.zeros(12), np.zeros(144)
    L.oracle_compute_bending(*[_d(_f64(a)) for a in (x0, x1, x2, x3, X0, X1, X2, X3)], beta, _d(W), _d(f), _d(K))
    return W[0], f, K.reshape(12, 12)


def compute_inertial(xa, xb, xc, Xa, Xb, Xc, g, rho):
    L = lib()
    L.oracle_compute_inertial.argtypes = [c_dp] * 7 + [ctypes.c_double] + [c_dp] * 3
    W, f, M = np.zeros(1), np.zeros(9), np.zeros(81)
    L.oracle_compute_inertial(*[_d(_f64(a)) for a in (xa, xb, xc, Xa, Xb, Xc, g)], rho, _d(W), _d(f), _d(M))
    return W[0], f, M.reshape(9, 9)


def face_frame(xa, xb, xc, Xa, Xb, Xc):
    PP, QQ = np.zeros(6), np.zeros(4)
    lib().oracle_face_frame(*[_d(_f64(a)) for a in (xa, xb, xc, Xa, Xb, Xc)], _d(PP), _d(QQ))
    return PP, QQ


# ---------------------------------------------------------------------------------------------------------------------
# Consumer of Forces::fill — numpy restatements, TEST INFRASTRUCTURE ONLY
# ---------------------------------------------------------------------------------------------------------------------
def _csc(mat, dof):
    import scipy.sparse as sp
    o, i, v = mat
    return sp.csc_matrix((v, i, o), shape=(dof, dof))


def cloth_rhs(M, f, v, h):
    """b = -(M v + h f): /root/reference/src/Cloth.cpp:345.  M = (outer, inner, values) as returned by forces_fill."""
    return -(_csc(M, f.size) @ v + h * f)


def eigen_cg(A, b, tol=np.finfo(float).eps, max_iter=None, fixed=None):
    """v = ConjugateGradient<SparseMatrix<double>, Lower|Upper>(A).solve(-b): /root/reference/src/GeneralizedSolver.cpp:122-125.
    Restates conjugate_gradient() of Eigen 3.3 (Eigen/src/IterativeLinearSolvers/ConjugateGradient.h — EXTERNAL dependency of the
    reference, not vendored under /root/reference; README.md:14 names Eigen 3.3, CMakeLists.txt:26 shows 3.3.4) with the default
    DiagonalPreconditioner (1/a_ii, 1 where a_ii == 0), x0 = 0, threshold tol^2 |rhs|^2, default cap 2 n iterations.
    Parity unpinned by the reference (the branch is dead code there); pinned here against a sparse direct solve (tests/test_oracle.py).
    fixed (bool per dof, optional): those dofs keep v = 0 and drop out of the system (the solve runs on the free-free block).
    Returns (v, iterations, relative residual)."""
    A = _csc(A, b.size) if isinstance(A, tuple) else A
    if fixed is not None and np.any(fixed):
        free = np.flatnonzero(~np.asarray(fixed, dtype=bool))
        sub = A.tocsr()[free][:, free].tocsc()
        xs, it, res = eigen_cg(sub, np.asarray(b, dtype=np.float64)[free], tol, max_iter)
        x = np.zeros(b.size)
        x[free] = xs
        return x, it, res
    rhs = -np.asarray(b, dtype=np.float64)
    n = rhs.size
    max_iter = 2 * n if max_iter is None else max_iter
    d = A.diagonal()
    dinv = np.where(d != 0.0, 1.0 / np.where(d != 0.0, d, 1.0), 1.0)
    x = np.zeros(n)
    residual = rhs.copy()
    rhs2 = float(rhs @ rhs)
    if rhs2 == 0.0:
        return x, 0, 0.0
    threshold = tol * tol * rhs2
    r2 = float(residual @ residual)
    if r2 < threshold:
        return x, 0, np.sqrt(r2 / rhs2)
    p = dinv * residual
    abs_new = float(residual @ p)
    it = 0
    while it < max_iter:
        tmp = A @ p
        alpha = abs_new / float(p @ tmp)
        x += alpha * p
        residual -= alpha * tmp
        r2 = float(residual @ residual)
        it += 1
        if r2 < threshold:
            break
        z = dinv * residual
        abs_old = abs_new
        abs_new = float(residual @ z)
        p = z + (abs_new / abs_old) * p
    return x, it, np.sqrt(r2 / rhs2)


def constraints_contact_rows(contacts, node_eol=None):
    """Contact part of Constraints::fill, /root/reference/src/Constraints.cpp:424-468: the (row, col, value) triplets Aineq_ receives for
    a CD2 contact list, row = running ineqsize; contacts touching an EoL node are skipped (`continue`, no row).  Returns a list of rows,
    each a list of (col, value) in push order."""
    rows = []
    for c in contacts:
        c1, c2 = int(c["count1"]), int(c["count2"])
        v2, w2 = c["verts2"], c["weights2"]
        if c1 == 3 and c2 == 1:                                              # :425-436
            if node_eol is not None and node_eol[v2[0]]:
                continue
            rows.append([(int(v2[0]) * 3 + k, -c["nor1"][k]) for k in range(3)])
        elif c1 == 2 and c2 == 2:                                            # :438-452
            if node_eol is not None and (node_eol[v2[0]] or node_eol[v2[1]]):
                continue
            rows.append([(int(v2[j]) * 3 + k, -c["nor2"][k] * w2[j]) for j in range(2) for k in range(3)])
        elif c1 == 1 and c2 == 3:                                            # :454-468
            if node_eol is not None and (node_eol[v2[0]] or node_eol[v2[1]] or node_eol[v2[2]]):
                continue
            rows.append([(int(v2[j]) * 3 + k, -c["nor2"][k] * w2[j]) for j in range(3) for k in range(3)])
    return rows


# ---------------------------------------------------------------------------------------------------------------------
# Per-step derived mesh data — numpy restatement, TEST INFRASTRUCTURE ONLY
# ---------------------------------------------------------------------------------------------------------------------
def constraints_fixed_rows(c, ci, v, eq_row0=0):
    """Restatement of the fixed-corner rows of Constraints::fill (Constraints.cpp:114-119, :470-497): c (4, 6) = mask + prescribed
    velocity of FixedList::c1..c4, ci (4,) node indices, v (N, 3) node velocities.  Returns (rows, cols, vals, beq) in push order."""
    c = np.asarray(c, float).reshape(4, 6)
    v = np.asarray(v, float).reshape(-1, 3)
    rows, cols, vals, beq = [], [], [], []
    eqsize = int(eq_row0)
    for k in range(4):
        if c[k, 0] != -1:
            for j in range(3):
                if c[k, j] == 1.0:
                    rows.append(eqsize); cols.append(int(ci[k]) * 3 + j); vals.append(c[k, j])
                    beq.append((1 - 0.01) * v[int(ci[k]), j] + c[k, j + 3])
                    eqsize += 1
    return np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals), np.array(beq)


def _arcsim_norm2(v):
    """vectors.hpp:108-109: dot() sums left to right starting from 0."""
    return (v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1]) + v[..., 2] * v[..., 2]


def _arcsim_cross(u, v):
    """vectors.hpp:113."""
    return np.stack([u[..., 1] * v[..., 2] - u[..., 2] * v[..., 1], u[..., 2] * v[..., 0] - u[..., 0] * v[..., 2],
                     u[..., 0] * v[..., 1] - u[..., 1] * v[..., 0]], axis=-1)


def _arcsim_normalize(v):
    """vectors.hpp:111: m == 0 ? 0 : u / m, and u / m is u * (1 / m) (vectors.hpp:104: one reciprocal, then products; the true-division
    specialisation at :130 is behind `#if defined(_AVX)`, which nothing defines).  Pinned by tests/test_forces_ref_pin.py against
    ArcSim's own compiled code."""
    m = np.sqrt(_arcsim_norm2(v))
    out = np.zeros_like(v)
    nz = m != 0
    out[nz] = (1.0 / m[nz])[:, None] * v[nz]
    return out


def mesh_normals(face_nodes, x):
    """compute_ws_data(Face*) (ArcSim mesh.cpp:135-140): face->n = normalize(cross(x1 - x0, x2 - x0)); normal<WS>(node)
    (geometry.cpp:302-316): n += cross(e1, e2) / (2 * norm2(e1) * norm2(e2)) over vert->adjf (faces in the order they were added
    = ascending face index), e1 / e2 to the next / next-but-one vertex of the face; then normalize.  Returns (face_n, node_n)."""
    fn = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    face_n = _arcsim_normalize(_arcsim_cross(x[fn[:, 1]] - x[fn[:, 0]], x[fn[:, 2]] - x[fn[:, 0]]))
    contrib = np.empty((fn.shape[0], 3, 3))
    for j in range(3):
        e1 = x[fn[:, (j + 1) % 3]] - x[fn[:, j]]
        e2 = x[fn[:, (j + 2) % 3]] - x[fn[:, j]]
        contrib[:, j] = (1.0 / ((2 * _arcsim_norm2(e1)) * _arcsim_norm2(e2)))[:, None] * _arcsim_cross(e1, e2)   # u / a = u * (1 / a)
    n = np.zeros_like(x)
    np.add.at(n, fn.reshape(-1), contrib.reshape(-1, 3))      # unbuffered, in index order: per node, faces ascending
    return face_n, _arcsim_normalize(n)


# ---- the reference's own Forces::fill (oracle/_ref/libforces_ref.so: Forces.cpp, UtilEOL.cpp, conversions.cpp, Compute*.cpp and ArcSim's
# mesh / geometry / util / vectors / transformation .cpp compiled UNMODIFIED against oracle/mini_eigen by oracle/Makefile; the prebuilt
# file travels to the GPU box).  adapter=True: libadapter_forces.so, the same driver with adapter/Forces_fill_b200.cpp in place of
# Forces.cpp — the drop-in body, executed on the GPU ------------------------------------------------------------------------------
def ref_forces_lib(adapter=False):
    """adapter: False = the reference's own Forces.cpp; True = adapter/Forces_fill_b200.cpp (results copied into the Eigen members);
    "zero_copy" = the same built with -DEOLC_ADAPTER_ZERO_COPY (the members' value arrays are the DMA targets)."""
    name = "libadapter_forces_zc.so" if adapter == "zero_copy" else "libadapter_forces.so" if adapter else "libforces_ref.so"
    if name in _REFLIBS:
        return _REFLIBS[name]
    path = os.path.join(_HERE, "_ref", name)
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and /root/reference is not here to build it from")
    L = ctypes.CDLL(path)
    L.ref_forces_fill.restype = ctypes.c_void_p
    L.ref_forces_fill.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, c_dp, c_ip, c_dp, c_dp, ctypes.c_double]
    L.ref_forces_new.restype = ctypes.c_void_p
    L.ref_forces_new.argtypes = L.ref_forces_fill.argtypes
    L.ref_forces_run.argtypes = [ctypes.c_void_p]
    L.ref_forces_time_steps.restype = ctypes.c_double
    L.ref_forces_time_steps.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double]
    L.ref_forces_mesh.restype = ctypes.c_void_p
    L.ref_forces_mesh.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, c_dp]
    L.ref_forces_free.argtypes = [ctypes.c_void_p]
    L.ref_forces_refill.argtypes = [ctypes.c_void_p, c_dp, c_dp]
    L.ref_forces_dof.argtypes = [ctypes.c_void_p]
    L.ref_forces_eol_cutoff.argtypes = [ctypes.c_void_p]
    L.ref_forces_f.restype = c_dp
    L.ref_forces_f.argtypes = [ctypes.c_void_p]
    L.ref_forces_nnz.restype = ctypes.c_int64
    L.ref_forces_nnz.argtypes = [ctypes.c_void_p, ctypes.c_int]
    for nm, rt in (("outer", c_ip), ("inner", c_ip), ("vals", c_dp)):
        fn = getattr(L, "ref_forces_" + nm)
        fn.restype = rt
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.ref_forces_n_edges.argtypes = [ctypes.c_void_p]
    L.ref_forces_edge_stencils.argtypes = [ctypes.c_void_p, c_ip]
    L.ref_forces_normals.argtypes = [ctypes.c_void_p, c_dp, c_dp, c_dp]
    _REFLIBS[name] = L
    return L


def _ref_forces_result(L, r):
    dof = L.ref_forces_dof(r)
    out = {"dof": dof, "EoL_cutoff": L.ref_forces_eol_cutoff(r), "f": np.ctypeslib.as_array(L.ref_forces_f(r), (dof,)).copy()}
    for which, name in ((0, "M"), (1, "MDK")):
        nnz = L.ref_forces_nnz(r, which)
        outer = np.ctypeslib.as_array(L.ref_forces_outer(r, which), (dof + 1,)).copy()
        inner = np.ctypeslib.as_array(L.ref_forces_inner(r, which), (max(nnz, 1),))[:nnz].copy()
        vals = np.ctypeslib.as_array(L.ref_forces_vals(r, which), (max(nnz, 1),))[:nnz].copy()
        out[name] = (outer, inner, vals)
    es = np.zeros((max(L.ref_forces_n_edges(r), 1), 4), np.int32)
    L.ref_forces_edge_stencils(r, _i(es))
    out["edge_stencil"] = es[:L.ref_forces_n_edges(r)]
    return out


def ref_forces_fill(face_nodes, x, X, mat=MATERIAL_DEFAULT, grav=GRAV_DEFAULT, h=H_DEFAULT, eol_index=None, adapter=False, more_steps=()):
    """The reference's Forces::fill (Forces.cpp:912-930), run as compiled from its own sources on a mesh built the way Cloth::build
    builds it (the edges come from ArcSim's Mesh::add(Face*)).  Returns dict(dof, EoL_cutoff, f, M=(outer, inner, vals), MDK=...,
    edge_stencil).  adapter=True: the same call reaches adapter/Forces_fill_b200.cpp instead (needs a CUDA device).
    more_steps: sequence of (x, X or None): Forces::fill is called again on the SAME Mesh / Forces objects after each update; the
    result is then a list of dicts (one per fill)."""
    L = ref_forces_lib(adapter)
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    X = _f64(X).reshape(-1, 2)
    N = x.shape[0]
    eol = None if eol_index is None else _i32(eol_index).reshape(N)
    r = L.ref_forces_fill(N, face_nodes.shape[0], _i(face_nodes), _d(x), _d(X), None if eol is None else _i(eol), _d(_f64(mat)),
                          _d(_f64(grav)), float(h))
    try:
        outs = [_ref_forces_result(L, r)]
        for xs, Xs in more_steps:
            xs = _f64(xs).reshape(N, 3)
            Xs = None if Xs is None else _f64(Xs).reshape(N, 2)
            L.ref_forces_refill(r, _d(xs), None if Xs is None else _d(Xs))
            outs.append(_ref_forces_result(L, r))
    finally:
        L.ref_forces_free(r)
    return outs if more_steps else outs[0]


def ref_forces_seconds(face_nodes, x, X, mat=MATERIAL_DEFAULT, grav=GRAV_DEFAULT, h=H_DEFAULT, instances=1, repeats=1, all_times=False):
    """Wall-clock seconds of Forces::fill ALONE (the meshes are built before the clock starts) as compiled from the reference's own
    sources: `instances` independent Mesh / Forces objects built from the same arrays, filled side by side on one thread each (the
    reference program is single-threaded; this is what a host with that many cores can get out of its code).  Forces::fill is
    called `repeats` times on the same objects (as Cloth::step does); returns the best time, or every time with all_times."""
    import threading
    L = ref_forces_lib()
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    X = _f64(X).reshape(-1, 2)
    m, g = _f64(mat), _f64(grav)
    hs = [L.ref_forces_new(x.shape[0], face_nodes.shape[0], _i(face_nodes), _d(x), _d(X), None, _d(m), _d(g), float(h)) for _ in range(instances)]
    times = []
    try:
        for _ in range(repeats):
            th = [threading.Thread(target=L.ref_forces_run, args=(hh,)) for hh in hs]     # ctypes releases the GIL for the call
            t = time.perf_counter()
            for t_ in th:
                t_.start()
            for t_ in th:
                t_.join()
            times.append(time.perf_counter() - t)
    finally:
        for hh in hs:
            L.ref_forces_free(hh)
    return times if all_times else min(times)


def ref_forces_step_seconds(face_nodes, x, X, mat=MATERIAL_DEFAULT, grav=GRAV_DEFAULT, h=H_DEFAULT, steps=5, adapter=False):
    """(seconds of the first Forces::fill, mean seconds of `steps` further fills on the same Mesh / Forces objects with every node
    moved a little in between) — the reference's own code, or with adapter=True the drop-in body on the GPU: flatten of the ArcSim
    pointer mesh + host -> device + kernels + device -> the members f / M / MDK of the reference's class Forces."""
    L = ref_forces_lib(adapter)
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    X = _f64(X).reshape(-1, 2)
    h_ = L.ref_forces_new(x.shape[0], face_nodes.shape[0], _i(face_nodes), _d(x), _d(X), None, _d(_f64(mat)), _d(_f64(grav)), float(h))
    try:
        t = time.perf_counter()
        L.ref_forces_run(h_)
        first = time.perf_counter() - t
        return first, L.ref_forces_time_steps(h_, int(steps), 1e-7)
    finally:
        L.ref_forces_free(h_)


def ref_mesh_data(face_nodes, x, X, x_new=None):
    """ArcSim's own mesh code on flat arrays: (edge_stencil in mesh.edges order, face normals, node normals) after compute_ws_data
    (mesh.cpp:135-151, geometry.cpp:302-316), optionally with new positions written to Node::x first."""
    L = ref_forces_lib()
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    X = _f64(X).reshape(-1, 2)
    N, F = x.shape[0], face_nodes.shape[0]
    r = L.ref_forces_mesh(N, F, _i(face_nodes), _d(x), _d(X))
    try:
        E = L.ref_forces_n_edges(r)
        es = np.zeros((max(E, 1), 4), np.int32)
        L.ref_forces_edge_stencils(r, _i(es))
        fn, nn = np.zeros((max(F, 1), 3)), np.zeros((max(N, 1), 3))
        xn = None if x_new is None else _f64(x_new).reshape(N, 3)
        L.ref_forces_normals(r, None if xn is None else _d(xn), _d(fn), _d(nn))
    finally:
        L.ref_forces_free(r)
    return es[:E], fn[:F], nn[:N]


# ---- the reference's own Constraints::fill (oracle/_ref/libconstraints_ref.so: Constraints.cpp + Collisions.cpp + boxTriCollision.cpp +
# Box.cpp + Obstacles.cpp + ... compiled UNMODIFIED, oracle/Makefile) ------------------------------------------------------------
def ref_constraints_fill(face_nodes, x, X, v, threshold, pxyz=None, pnorms=None, box_whd=None, box_E=None, fixed_c=None, fixed_ci=None,
                         corner_id=None, h=H_DEFAULT):
    """Constraints::fill (Constraints.cpp:122-513) as compiled from the reference's own sources: CD2 inside (:423), the contact rows
    (:424-468), the fixed-corner rows (:470-497).  Returns dict(Aineq=(n_rows, outer, inner, vals) [column-major compressed, as
    Eigen stores it], bineq, Aeq=..., beq, hasFixed, hasCollisions).  corner_id: see oracle/ref_constraints_driver.cpp."""
    name = "libconstraints_ref.so"
    if name not in _REFLIBS:
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing and /root/reference is not here to build it from")
        L = ctypes.CDLL(path)
        L.ref_constraints_fill.restype = ctypes.c_void_p
        L.ref_constraints_fill.argtypes = [ctypes.c_int, ctypes.c_int, c_ip, c_dp, c_dp, c_dp, c_ip, ctypes.c_double, ctypes.c_int, c_dp, c_dp,
                                           ctypes.c_int, c_dp, c_dp, c_dp, c_ip, ctypes.c_double]
        L.ref_constraints_free.argtypes = [ctypes.c_void_p]
        for nm, rt in (("rows", ctypes.c_int), ("cols", ctypes.c_int), ("nnz", ctypes.c_int64), ("outer", c_ip), ("inner", c_ip), ("vals", c_dp), ("b", c_dp)):
            fn = getattr(L, "ref_constraints_" + nm)
            fn.restype = rt
            fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ref_constraints_flags.argtypes = [ctypes.c_void_p]
        _REFLIBS[name] = L
    L = _REFLIBS[name]
    face_nodes = _i32(face_nodes).reshape(-1, 3)
    x = _f64(x).reshape(-1, 3)
    X = _f64(X).reshape(-1, 2)
    v = _f64(v).reshape(-1, 3)
    N = x.shape[0]
    pxyz = _f64(np.zeros((0, 3)) if pxyz is None else pxyz).reshape(-1, 3)
    pnorms = _f64(np.zeros((0, 3)) if pnorms is None else pnorms).reshape(-1, 3)
    box_whd = _f64(np.zeros((0, 3)) if box_whd is None else box_whd).reshape(-1, 3)
    box_E = _f64(np.zeros((0, 16)) if box_E is None else box_E).reshape(-1, 16)
    if box_whd.shape[0] and not ref_hash_defined(N, face_nodes.shape[0]):
        raise ValueError("the reference's int edge hash overflows (undefined behaviour) for this mesh size")
    fc = _f64(-np.ones((4, 6)) if fixed_c is None else fixed_c).reshape(4, 6)
    fi = _i32(np.zeros(4) if fixed_ci is None else fixed_ci).reshape(4)
    cid = None if corner_id is None else _i32(corner_id).reshape(N)
    r = L.ref_constraints_fill(N, face_nodes.shape[0], _i(face_nodes), _d(x), _d(X), _d(v), None if cid is None else _i(cid), float(threshold),
                               pxyz.shape[0], _d(pxyz), _d(pnorms), box_whd.shape[0], _d(box_whd), _d(box_E), _d(fc), _i(fi), float(h))
    try:
        out = {}
        for which, nm, bn in ((0, "Aineq", "bineq"), (1, "Aeq", "beq")):
            rows, cols, nnz = L.ref_constraints_rows(r, which), L.ref_constraints_cols(r, which), L.ref_constraints_nnz(r, which)
            outer = np.ctypeslib.as_array(L.ref_constraints_outer(r, which), (cols + 1,)).copy()
            inner = np.ctypeslib.as_array(L.ref_constraints_inner(r, which), (nnz,)).copy() if nnz else np.zeros(0, np.int32)
            vals = np.ctypeslib.as_array(L.ref_constraints_vals(r, which), (nnz,)).copy() if nnz else np.zeros(0)
            out[nm] = (rows, outer, inner, vals)
            out[bn] = np.ctypeslib.as_array(L.ref_constraints_b(r, which), (rows,)).copy() if rows else np.zeros(0)
        fl = L.ref_constraints_flags(r)
        out["hasFixed"], out["hasCollisions"] = bool(fl & 1), bool(fl & 2)
    finally:
        L.ref_constraints_free(r)
    return out
