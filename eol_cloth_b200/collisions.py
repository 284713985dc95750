"""Host-side mirror of the reference's CD / CD2 (/root/reference/src/Collisions.h:9-11) over the C ABI."""
import ctypes
from collections import namedtuple

import numpy as np

from . import capi

# POD mirror of btc::Collision (include/eolc.h eolc_contact)
CONTACT_DTYPE = np.dtype([
    ("dist", "f8"), ("nor1", "f8", 3), ("nor2", "f8", 3), ("pos1", "f8", 3), ("pos2", "f8", 3), ("pos1_", "f8", 3),
    ("weights1", "f8", 3), ("weights2", "f8", 3), ("edgeDir", "f8", 3),
    ("count1", "i4"), ("count2", "i4"), ("verts1", "i4", 3), ("verts2", "i4", 3), ("tri1", "i4"), ("tri2", "i4"),
    ("edge1", "i4", 3), ("n_edge1", "i4"), ("edge2", "i4"), ("reserved", "i4"),
])
assert CONTACT_DTYPE.itemsize == 264

# what CD/CD2 read from `Obstacles` (Obstacles.h, Points.h, Box.h): cdthreshold, points->pxyz/norms, boxes[b]->dim/E1
Obstacles = namedtuple("Obstacles", "cdthreshold pxyz pnorms box_whd box_E")


def make_obstacles(cdthreshold, pxyz=None, pnorms=None, box_whd=None, box_E=None):
    z3 = np.zeros((0, 3))
    return Obstacles(float(cdthreshold),
                     capi.f64(z3 if pxyz is None else pxyz).reshape(-1, 3),
                     capi.f64(z3 if pnorms is None else pnorms).reshape(-1, 3),
                     capi.f64(z3 if box_whd is None else box_whd).reshape(-1, 3),
                     capi.f64(np.zeros((0, 16)) if box_E is None else box_E).reshape(-1, 16))


class CollisionPlan:
    """eolc_cd_plan: btc edge table (createEdges order) + perturbation table for (N, threshold)."""

    def __init__(self, ctx, n_nodes, face_nodes, threshold):
        self.ctx = ctx
        self._h = capi.c_vp()
        fn = capi.i32(face_nodes).reshape(-1, 3)
        capi.check(capi.lib().eolc_cd_plan_create(ctx.handle, int(n_nodes), fn.shape[0], capi.iptr(fn), float(threshold),
                                                  ctypes.byref(self._h)))
        self.N, self.F, self.threshold = int(n_nodes), fn.shape[0], float(threshold)
        self.E = capi.lib().eolc_cd_edge_count(self._h)

    def edge_table(self):
        out = np.empty((max(self.E, 1), 6), dtype=np.int32)
        capi.check(capi.lib().eolc_cd_edge_table(self._h, capi.iptr(out)))
        return out[:self.E].copy()

    def run(self, x, obs, point_eol_flag, remap, capacity=None, x_is_device_ptr=False, n_scenes=1, out=None):
        """out: optional caller-owned record buffer (CONTACT_DTYPE), reused across calls; page-locked memory is DMA'd into directly."""
        if out is not None:
            capacity = out.shape[0]
        if capacity is None:
            capacity = n_scenes * (self.N + 64) + 4096
        nP, nB = obs.pxyz.shape[0], obs.box_whd.shape[0]
        L = capi.lib()
        caller_out = out
        while True:
            out = caller_out if caller_out is not None else np.empty(capacity, dtype=CONTACT_DTYPE)
            if x_is_device_ptr:
                off = np.zeros(n_scenes + 1, dtype=np.int32)
                rc = L.eolc_cd_run_batched_dev(self._h, int(n_scenes), x, nP, capi.dptr(obs.pxyz), capi.dptr(obs.pnorms), nB,
                                               capi.dptr(obs.box_whd), capi.dptr(obs.box_E), int(point_eol_flag), int(remap),
                                               out.ctypes.data_as(capi.c_vp), capacity, capi.iptr(off))
                n = int(off[-1])
            else:
                xx = capi.f64(x).reshape(-1)
                if xx.size != 3 * self.N:
                    raise capi.EolcError("x must hold 3N doubles")
                nn = ctypes.c_int32(0)
                rc = L.eolc_cd_run(self._h, capi.dptr(xx), nP, capi.dptr(obs.pxyz), capi.dptr(obs.pnorms), nB,
                                   capi.dptr(obs.box_whd), capi.dptr(obs.box_E), int(point_eol_flag), int(remap),
                                   out.ctypes.data_as(capi.c_vp), capacity, ctypes.byref(nn))
                n, off = nn.value, None
            if rc == -3 and caller_out is None:       # EOLC_ERR_CAPACITY: n holds the required size
                capacity = n
                continue
            capi.check(rc)
            return (out[:n], off) if x_is_device_ptr else out[:n]

    def run_resident(self, x_ptr, obs, point_eol_flag, remap, n_scenes=1):
        """Device-resident run: x_ptr is a device pointer to n_scenes states; the records stay on the device (contacts_dev()),
        only the per-scene offsets (n_scenes + 1) come back."""
        off = np.zeros(n_scenes + 1, dtype=np.int32)
        nP, nB = obs.pxyz.shape[0], obs.box_whd.shape[0]
        capi.check(capi.lib().eolc_cd_run_batched_resident_dev(self._h, int(n_scenes), x_ptr, nP, capi.dptr(obs.pxyz), capi.dptr(obs.pnorms),
                                                               nB, capi.dptr(obs.box_whd), capi.dptr(obs.box_E), int(point_eol_flag),
                                                               int(remap), capi.iptr(off)))
        return off

    def contacts_dev(self):
        """(device pointer, count) of the records of the last run."""
        p, n = capi.c_vp(), ctypes.c_int32(0)
        capi.check(capi.lib().eolc_cd_contacts_dev(self._h, ctypes.byref(p), ctypes.byref(n)))
        return p.value, n.value

    def _row_buffers(self, cap):
        """Page-locked result arrays kept across calls (allocating and first-touching 20 MB of fresh pageable memory per call cost
        more than the device work and the copy together)."""
        b = getattr(self, "_rowbuf", None)
        if b is None or b[0] < cap:
            for hb in (b[1:] if b else ()):
                hb.free()
            c = max(cap, 1) + max(cap, 1) // 4
            b = self._rowbuf = (c, capi.HostBuffer(c + 1, np.int32), capi.HostBuffer(9 * c, np.int32), capi.HostBuffer(9 * c, np.float64))
        return b

    def contact_rows(self, node_eol=None):
        """Inequality rows of the contacts of the LAST run, built on the device (Constraints.cpp:424-468): (row_nnz, cols, vals) with
        9 slots per row in the reference's triplet order.  The arrays are views of buffers owned by the plan, valid until its
        next contact_rows call."""
        cap, nnz, cols, vals = self._row_buffers(int(capi.lib().eolc_cd_last_count(self._h)))
        n = ctypes.c_int32(0)
        eol = None if node_eol is None else np.ascontiguousarray(node_eol, dtype=np.uint8)
        capi.check(capi.lib().eolc_cd_contact_rows(self._h, None if eol is None else eol.ctypes.data_as(capi.c_vp), cap, ctypes.byref(n),
                                                   capi.iptr(nnz.array), capi.iptr(cols.array), capi.dptr(vals.array)))
        return nnz.array[:n.value], cols.array[:9 * n.value].reshape(-1, 9), vals.array[:9 * n.value].reshape(-1, 9)

    def contact_rows_csr(self, node_eol=None):
        """The same rows compacted on the device: (row_ptr, cols, vals) — Aineq's triplets are (r, cols[q], vals[q]) for
        row_ptr[r] <= q < row_ptr[r + 1], in the reference's push order.  Views of buffers owned by the plan."""
        cap, rp, cols, vals = self._row_buffers(int(capi.lib().eolc_cd_last_count(self._h)))
        n, nz = ctypes.c_int32(0), ctypes.c_int32(0)
        eol = None if node_eol is None else np.ascontiguousarray(node_eol, dtype=np.uint8)
        capi.check(capi.lib().eolc_cd_contact_rows_csr(self._h, None if eol is None else eol.ctypes.data_as(capi.c_vp), cap, 9 * cap,
                                                       ctypes.byref(n), ctypes.byref(nz), capi.iptr(rp.array), capi.iptr(cols.array),
                                                       capi.dptr(vals.array)))
        return rp.array[:n.value + 1], cols.array[:nz.value], vals.array[:nz.value]

    def stats(self):
        pt, ln = ctypes.c_int64(), ctypes.c_int32()
        capi.check(capi.lib().eolc_cd_last_stats(self._h, ctypes.byref(pt), ctypes.byref(ln)))
        return pt.value, ln.value

    def close(self):
        if self._h:
            for hb in (getattr(self, "_rowbuf", None) or ())[1:]:
                hb.free()
            self._rowbuf = None
            capi.lib().eolc_cd_plan_destroy(self._h)
            self._h = capi.c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_PLANS = {}


def _plan_for(ctx, mesh, obs):
    fn = capi.i32(mesh["face_nodes"]).reshape(-1, 3)
    N = np.asarray(mesh["x"]).reshape(-1, 3).shape[0]
    key = (id(ctx), N, fn.shape[0], hash(fn.tobytes()), obs.cdthreshold)
    p = _PLANS.get(key)
    if p is None:
        if len(_PLANS) > 8:
            _PLANS.clear()
        p = _PLANS[key] = CollisionPlan(ctx, N, fn, obs.cdthreshold)
    return p


def CD(ctx, mesh, obs, cls):
    """void CD(const Mesh&, shared_ptr<Obstacles>, vector<shared_ptr<btc::Collision>>& cls) — Collisions.cpp:11-53.
    Appends to the list `cls` (records of dtype CONTACT_DTYPE)."""
    cls.extend(_plan_for(ctx, mesh, obs).run(mesh["x"], obs, point_eol_flag=1, remap=1))


def CD2(ctx, mesh, obs, cls):
    """void CD2(...) — Collisions.cpp:55-78."""
    cls.extend(_plan_for(ctx, mesh, obs).run(mesh["x"], obs, point_eol_flag=0, remap=0))


def contact_rows(contacts, node_eol=None):
    """Constraints::fill, contact part (Constraints.cpp:424-468) for a host contact list: (row_nnz, cols, vals), 9 slots per row."""
    c = np.ascontiguousarray(contacts, dtype=CONTACT_DTYPE)
    cap = max(len(c), 1)
    nnz, cols, vals = np.zeros(cap, np.int32), np.full(9 * cap, -1, np.int32), np.zeros(9 * cap)
    n = ctypes.c_int32(0)
    eol = None if node_eol is None else np.ascontiguousarray(node_eol, dtype=np.uint8)
    capi.check(capi.lib().eolc_constraints_contact_rows(c.ctypes.data_as(capi.c_vp), len(c), None if eol is None else eol.ctypes.data_as(capi.c_vp),
                                                        ctypes.byref(n), capi.iptr(nnz), capi.iptr(cols), capi.dptr(vals)))
    return nnz[:n.value], cols[:9 * n.value].reshape(-1, 9), vals[:9 * n.value].reshape(-1, 9)
