"""ctypes binding of include/eolc.h (libeolc_b200.so).  Fails loudly if the CUDA library is missing."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int32)
c_vp = ctypes.c_void_p


class EolcError(RuntimeError):
    pass


class MaterialC(ctypes.Structure):
    # struct Material, /root/reference/src/Cloth.h:28-35
    _fields_ = [("density", ctypes.c_double), ("e", ctypes.c_double), ("nu", ctypes.c_double),
                ("beta", ctypes.c_double), ("dampingA", ctypes.c_double), ("dampingB", ctypes.c_double)]


def lib_path():
    # EOLC_LIB: developer knob to A/B an alternative build of the same library (scripts/gpu_variants.sh)
    return os.environ.get("EOLC_LIB") or os.path.join(_HERE, "libeolc_b200.so")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise EolcError(f"{path} is missing: build it with `make -C eol_cloth_b200/csrc` "
                        "(or python -c 'import __graft_entry__ as g; g.build()'). There is no CPU fallback.")
    L = ctypes.CDLL(path)
    L.eolc_last_error.restype = ctypes.c_char_p
    L.eolc_ctx_create.argtypes = [ctypes.c_int, ctypes.POINTER(c_vp)]
    L.eolc_ctx_destroy.argtypes = [c_vp]
    L.eolc_ctx_stream.restype = c_vp
    L.eolc_ctx_stream.argtypes = [c_vp]
    L.eolc_mesh_edge_stencils.argtypes = [ctypes.c_int32, ctypes.c_int32, c_ip, c_ip, c_ip]
    L.eolc_forces_plan_create.argtypes = [c_vp, ctypes.c_int32, ctypes.c_int32, c_ip, ctypes.c_int32, c_ip, c_ip, c_dp,
                                          ctypes.POINTER(c_vp)]
    L.eolc_forces_plan_destroy.argtypes = [c_vp]
    L.eolc_forces_pattern.argtypes = [c_vp, ctypes.c_int, c_ip, ctypes.POINTER(ctypes.c_int64),
                                      ctypes.POINTER(c_ip), ctypes.POINTER(c_ip)]
    L.eolc_forces_counts.argtypes = [c_vp, c_ip, c_ip]
    L.eolc_forces_fill.argtypes = [c_vp, c_dp, c_dp, ctypes.POINTER(MaterialC), c_dp, ctypes.c_double, c_dp, c_dp, c_dp]
    L.eolc_forces_fill_dev.argtypes = [c_vp, c_vp, c_vp, ctypes.POINTER(MaterialC), c_dp, ctypes.c_double, c_vp, c_vp, c_vp]
    L.eolc_forces_fill_batched_dev.argtypes = [c_vp, ctypes.c_int32, c_vp, c_vp, ctypes.POINTER(MaterialC), c_dp,
                                               ctypes.c_double, c_vp, c_vp, c_vp]
    L.eolc_forces_fill_ex.argtypes = [c_vp, c_dp, c_dp, ctypes.POINTER(MaterialC), c_dp, ctypes.c_double, c_dp, c_dp, c_dp, ctypes.c_uint32]
    L.eolc_forces_fill_batched_dev_ex.argtypes = [c_vp, ctypes.c_int32, c_vp, c_vp, ctypes.POINTER(MaterialC), c_dp,
                                                  ctypes.c_double, c_vp, c_vp, c_vp, ctypes.c_uint32]
    L.eolc_host_alloc.restype = c_vp
    L.eolc_host_alloc.argtypes = [ctypes.c_size_t]
    L.eolc_host_free.argtypes = [c_vp]
    L.eolc_forces_launches_per_fill.argtypes = [c_vp]
    L.eolc_mesh_normals.argtypes = [c_vp, c_dp, c_dp, c_dp]
    L.eolc_mesh_normals_dev.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.eolc_forces_rhs_dev.argtypes = [c_vp, c_vp, c_vp, c_vp, ctypes.c_double, c_vp]
    L.eolc_forces_integrate_dev.argtypes = [c_vp, c_vp, ctypes.c_double, c_vp]
    L.eolc_forces_integrate_X_dev.argtypes = [c_vp, c_vp, ctypes.c_double, c_vp]
    L.eolc_solve_cg_dev.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_double, ctypes.c_int32, c_ip, c_dp]
    L.eolc_cd_plan_create.argtypes = [c_vp, ctypes.c_int32, ctypes.c_int32, c_ip, ctypes.c_double, ctypes.POINTER(c_vp)]
    L.eolc_cd_plan_destroy.argtypes = [c_vp]
    L.eolc_cd_edge_count.argtypes = [c_vp]
    L.eolc_cd_edge_table.argtypes = [c_vp, c_ip]
    L.eolc_cd_run.argtypes = [c_vp, c_dp, ctypes.c_int32, c_dp, c_dp, ctypes.c_int32, c_dp, c_dp, ctypes.c_int,
                              ctypes.c_int, c_vp, ctypes.c_int32, c_ip]
    L.eolc_cd_run_dev.argtypes = [c_vp, c_vp, ctypes.c_int32, c_dp, c_dp, ctypes.c_int32, c_dp, c_dp, ctypes.c_int,
                                  ctypes.c_int, c_vp, ctypes.c_int32, c_ip]
    L.eolc_cd_run_batched_dev.argtypes = [c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_dp, c_dp, ctypes.c_int32, c_dp,
                                          c_dp, ctypes.c_int, ctypes.c_int, c_vp, ctypes.c_int32, c_ip]
    L.eolc_cd_run_batched_resident_dev.argtypes = [c_vp, ctypes.c_int32, c_vp, ctypes.c_int32, c_dp, c_dp, ctypes.c_int32, c_dp,
                                                   c_dp, ctypes.c_int, ctypes.c_int, c_ip]
    L.eolc_cd_contacts_dev.argtypes = [c_vp, ctypes.POINTER(c_vp), c_ip]
    L.eolc_cd_angle_cuts.argtypes = [c_dp, c_dp, c_dp]
    L.eolc_constraints_contact_rows.argtypes = [c_vp, ctypes.c_int32, c_vp, c_ip, c_ip, c_ip, c_dp]
    L.eolc_constraints_fixed_rows.argtypes = [c_dp, c_ip, c_dp, ctypes.c_int32, ctypes.c_int32, c_ip, c_ip, c_ip, c_dp, c_dp]
    L.eolc_cd_contact_rows.argtypes = [c_vp, c_vp, ctypes.c_int32, c_ip, c_ip, c_ip, c_dp]
    L.eolc_cd_contact_rows_csr.argtypes = [c_vp, c_vp, ctypes.c_int32, ctypes.c_int32, c_ip, c_ip, c_ip, c_ip, c_dp]
    L.eolc_cd_last_count.argtypes = [c_vp]
    L.eolc_cd_last_stats.argtypes = [c_vp, ctypes.POINTER(ctypes.c_int64), c_ip]
    _LIB = L
    return L


FILL_M_UNCHANGED = 1   # EOLC_FILL_M_UNCHANGED
FILL_EXACT_SYMMETRY = 2   # EOLC_FILL_EXACT_SYMMETRY


class HostBuffer:
    """Page-locked host array from eolc_host_alloc (the DMA target itself; pageable arrays cost a staging copy)."""

    def __init__(self, shape, dtype=np.float64):
        self.array = None
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self._p = lib().eolc_host_alloc(max(n, 1))
        if not self._p:
            raise EolcError("eolc_host_alloc: " + lib().eolc_last_error().decode())
        buf = (ctypes.c_char * max(n, 1)).from_address(self._p)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if getattr(self, "_p", None):
            self.array = None
            lib().eolc_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def check(rc):
    if rc != 0:
        raise EolcError(f"eolc error {rc}: {lib().eolc_last_error().decode()}")


def device_count():
    return lib().eolc_device_count()


def dptr(a):
    return a.ctypes.data_as(c_dp)


def iptr(a):
    return a.ctypes.data_as(c_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """eolc_ctx: one CUDA device + stream. Raises EolcError when no CUDA device is present (no CPU fallback)."""

    def __init__(self, device=0):
        self._h = c_vp()
        check(lib().eolc_ctx_create(int(device), ctypes.byref(self._h)))
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    @property
    def stream(self):
        return lib().eolc_ctx_stream(self._h)

    def close(self):
        if self._h:
            lib().eolc_ctx_destroy(self._h)
            self._h = c_vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
