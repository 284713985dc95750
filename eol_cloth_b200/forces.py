"""Host-side mirror of the reference's `Forces` class (/root/reference/src/Forces.h:26-45) over the C ABI.

    forces = Forces(ctx)
    forces.fill(mesh_arrays, material, grav, h)      # == Forces::fill(mesh, mat, grav, h), Forces.cpp:912-930
    forces.f, forces.M, forces.MDK, forces.EoL_cutoff

`M` / `MDK` are (outer, inner, values) triples laid out exactly like the reference's column-major
Eigen::SparseMatrix<double> (outerIndexPtr / innerIndexPtr / valuePtr).
"""
import ctypes
from collections import namedtuple

import numpy as np

from . import capi

Material = namedtuple("Material", "density e nu beta dampingA dampingB")
# simulationSettings.json:27-33
Material.DEFAULT = Material(0.05, 50.0, 0.01, 1.0e-5, 0.0, 1.0)


def _matc(mat):
    return capi.MaterialC(*[float(v) for v in mat])


class ForcesPlan:
    """eolc_forces_plan: the fixed CSR pattern + element->slot maps; rebuild only after a remesh."""

    def __init__(self, ctx, n_nodes, face_nodes, edge_stencil, eol_index=None, X_hint=None):
        self.ctx = ctx
        self._h = capi.c_vp()
        fn = capi.i32(face_nodes).reshape(-1, 3)
        es = capi.i32(edge_stencil).reshape(-1, 4)
        eol = None if eol_index is None else capi.i32(eol_index)
        xh = None if X_hint is None else capi.f64(X_hint)
        capi.check(capi.lib().eolc_forces_plan_create(
            ctx.handle, int(n_nodes), fn.shape[0], capi.iptr(fn), es.shape[0], capi.iptr(es),
            None if eol is None else capi.iptr(eol), None if xh is None else capi.dptr(xh), ctypes.byref(self._h)))
        self.N = int(n_nodes)
        dof = ctypes.c_int32()
        capi.check(capi.lib().eolc_forces_pattern(self._h, 0, ctypes.byref(dof), None, None, None))
        self.dof = dof.value           # 3N + 2 EoL_Count (Forces.cpp:914)
        nf, ne = ctypes.c_int32(), ctypes.c_int32()
        capi.check(capi.lib().eolc_forces_counts(self._h, ctypes.byref(nf), ctypes.byref(ne)))
        self.n_faces, self.n_interior_edges = nf.value, ne.value
        self.nnz = [self._nnz(0), self._nnz(1)]

    @property
    def handle(self):
        return self._h

    def _nnz(self, which):
        nnz = ctypes.c_int64()
        capi.check(capi.lib().eolc_forces_pattern(self._h, which, None, ctypes.byref(nnz), None, None))
        return nnz.value

    def pattern(self, which):
        """(outer, inner) int32 arrays == Eigen outerIndexPtr()/innerIndexPtr() of M (0) or MDK (1)."""
        dof, nnz = ctypes.c_int32(), ctypes.c_int64()
        outer, inner = capi.c_ip(), capi.c_ip()
        capi.check(capi.lib().eolc_forces_pattern(self._h, which, ctypes.byref(dof), ctypes.byref(nnz),
                                                  ctypes.byref(outer), ctypes.byref(inner)))
        o = np.ctypeslib.as_array(outer, (dof.value + 1,)).copy()
        i = np.ctypeslib.as_array(inner, (nnz.value,)).copy() if nnz.value else np.zeros(0, np.int32)
        return o, i

    @property
    def launches_per_fill(self):
        return capi.lib().eolc_forces_launches_per_fill(self._h)

    def fill(self, x, X, mat, grav, h):
        """Host arrays in, host arrays out (H2D / D2H inside): returns f, M_vals, MDK_vals."""
        x = capi.f64(x).reshape(-1)
        X = capi.f64(X).reshape(-1)
        if x.size != 3 * self.N or X.size != 2 * self.N:
            raise capi.EolcError("x must hold 3N and X 2N doubles")
        g = capi.f64(grav)
        f = np.empty(self.dof)
        Mv = np.empty(self.nnz[0])
        Kv = np.empty(self.nnz[1])
        m = _matc(mat)
        capi.check(capi.lib().eolc_forces_fill(self._h, capi.dptr(x), capi.dptr(X), ctypes.byref(m), capi.dptr(g),
                                               float(h), capi.dptr(f), capi.dptr(Mv), capi.dptr(Kv)))
        return f, Mv, Kv

    def fill_into(self, x, X, mat, grav, h, f, Mv, Kv, m_unchanged=False):
        """Same as fill() but into caller-provided contiguous float64 arrays (no allocation in the timed region).
        m_unchanged: EOLC_FILL_M_UNCHANGED — X and the density are those of the previous fill, Mv is left as it is."""
        g = capi.f64(grav)
        m = _matc(mat)
        capi.check(capi.lib().eolc_forces_fill_ex(self._h, capi.dptr(x), capi.dptr(X), ctypes.byref(m), capi.dptr(g),
                                                  float(h), capi.dptr(f), capi.dptr(Mv), capi.dptr(Kv),
                                                  capi.FILL_M_UNCHANGED if m_unchanged else 0))

    def fill_dev(self, x_ptr, X_ptr, mat, grav, h, f_ptr, Mv_ptr, Kv_ptr, n_scenes=1, m_unchanged=False, exact_symmetry=False):
        """Device pointers (ints, e.g. torch.Tensor.data_ptr()); asynchronous on ctx.stream.
        exact_symmetry: EOLC_FILL_EXACT_SYMMETRY — MDK symmetric bit for bit, as the host entries always deliver it."""
        g = capi.f64(grav)
        m = _matc(mat)
        capi.check(capi.lib().eolc_forces_fill_batched_dev_ex(self._h, int(n_scenes), x_ptr, X_ptr, ctypes.byref(m),
                                                              capi.dptr(g), float(h), f_ptr, Mv_ptr, Kv_ptr,
                                                              (capi.FILL_M_UNCHANGED if m_unchanged else 0) |
                                                              (capi.FILL_EXACT_SYMMETRY if exact_symmetry else 0)))

    def rhs_dev(self, Mv_ptr, f_ptr, v_ptr, h, b_ptr):
        """b = -(M v + h f) on the device (Cloth::solve, Cloth.cpp:345); device pointers, asynchronous on ctx.stream."""
        capi.check(capi.lib().eolc_forces_rhs_dev(self._h, Mv_ptr, f_ptr, v_ptr, float(h), b_ptr))

    def integrate_dev(self, v_ptr, h, x_ptr):
        """x += h v on the device (Cloth::step, Cloth.cpp:394-400)."""
        capi.check(capi.lib().eolc_forces_integrate_dev(self._h, v_ptr, float(h), x_ptr))

    def integrate_X_dev(self, v_ptr, h, X_ptr):
        """X += h v_X for the EoL nodes on the device (Cloth::step, Cloth.cpp:401-407); no-op without EoL nodes."""
        capi.check(capi.lib().eolc_forces_integrate_X_dev(self._h, v_ptr, float(h), X_ptr))

    def solve_cg_dev(self, Kv_ptr, b_ptr, v_ptr, tol=2.220446049250313e-16, max_iter=None, fixed_ptr=None):
        """v = ConjugateGradient(MDK).solve(-b) on the device (GeneralizedSolver.cpp:120-126, Eigen defaults: diagonal
        preconditioner, x0 = 0, tol = epsilon, at most 2 dof iterations).  fixed_ptr: optional device pointer to dof bytes, non-zero =
        the dof keeps velocity 0.  Returns (iterations issued, relative residual)."""
        it, res = ctypes.c_int32(0), ctypes.c_double(0.0)
        capi.check(capi.lib().eolc_solve_cg_dev(self._h, Kv_ptr, b_ptr, fixed_ptr, v_ptr, float(tol), int(2 * self.dof if max_iter is None else max_iter),
                                                ctypes.byref(it), ctypes.byref(res)))
        return it.value, res.value

    def normals(self, x):
        """World-space face and node normals of compute_ws_data (ArcSim mesh.cpp:135-143, geometry.cpp:302-316) for the state x:
        returns (face_n (F,3), node_n (N,3)).  Host arrays in and out."""
        x = capi.f64(x).reshape(-1)
        if x.size != 3 * self.N:
            raise capi.EolcError("x must hold 3N doubles")
        fnorm = np.empty((self.n_faces, 3))
        nnorm = np.empty((self.N, 3))
        capi.check(capi.lib().eolc_mesh_normals(self._h, capi.dptr(x), capi.dptr(fnorm) if self.n_faces else None, capi.dptr(nnorm)))
        return fnorm, nnorm

    def normals_dev(self, x_ptr, face_n_ptr, node_n_ptr):
        """Same on device pointers (either output may be None); asynchronous on ctx.stream."""
        capi.check(capi.lib().eolc_mesh_normals_dev(self._h, x_ptr, face_n_ptr, node_n_ptr))

    def close(self):
        if self._h:
            capi.lib().eolc_forces_plan_destroy(self._h)
            self._h = capi.c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Forces:
    """Mirror of `class Forces` (Forces.h:26-45): members f, M, MDK, EoL_cutoff; method fill()."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.f = None
        self.M = None
        self.MDK = None
        self.EoL_cutoff = 0
        self._plan = None
        self._topo_key = None

    def fill(self, mesh, mat, grav, h):
        """mesh: dict(x (N,3), X (N,2), face_nodes (F,3), edge_stencil (E,4)[, eol_index (N,): Node::EoL_index, -1 = Lagrangian]) —
        the flattened ArcSim mesh.  The topology plan is cached and rebuilt when the topology arrays change (remesh)."""
        fn = capi.i32(mesh["face_nodes"]).reshape(-1, 3)
        es = capi.i32(mesh["edge_stencil"]).reshape(-1, 4)
        N = np.asarray(mesh["x"]).reshape(-1, 3).shape[0]
        eol = mesh.get("eol_index")
        eol = None if eol is None else capi.i32(eol)
        key = (N, fn.shape[0], es.shape[0], hash(fn.tobytes()), hash(es.tobytes()), None if eol is None else hash(eol.tobytes()))
        if key != self._topo_key:
            if self._plan is not None:
                self._plan.close()
            self._plan = ForcesPlan(self.ctx, N, fn, es, eol, mesh.get("X"))
            self._topo_key = key
            self._pat = [self._plan.pattern(0), self._plan.pattern(1)]
        f, Mv, Kv = self._plan.fill(mesh["x"], mesh["X"], mat, grav, h)
        self.f = f
        self.M = (self._pat[0][0], self._pat[0][1], Mv)
        self.MDK = (self._pat[1][0], self._pat[1][1], Kv)
        self.EoL_cutoff = 3 * N           # Forces.cpp:919
        return self
