"""Synthetic cloth meshes and states for the BASELINE configs, flat arrays as the C ABI expects.

regular2(n): n x n grid, 2 triangles per cell (the "1024x1024 sheet (~2M triangles)" of BASELINE.json).
build4(n, m): the mesh Cloth::build makes (/root/reference/src/Cloth.cpp:36-148): grid nodes (i, j) -> i*m + j, then one
centre node per cell, four faces per cell (k0,k0+1,kc) (k0+1,k0+m+1,kc) (k0+m+1,k0+m,kc) (k0+m,k0,kc).
edge stencils follow ArcSim's Mesh::add(Face) edge creation (mesh.cpp:356-378) via eolc_mesh_edge_stencils.
"""
import ctypes

import numpy as np

from . import capi


def regular2(n, m=None):
    """Returns (X (N,2), face_nodes (F,3) int32). Node (i,j) -> i*m + j, X = (i/(n-1), j/(m-1))."""
    m = n if m is None else m
    i, j = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    X = np.stack([i.ravel() / (n - 1), j.ravel() / (m - 1)], axis=1).astype(np.float64)
    ci, cj = np.meshgrid(np.arange(n - 1), np.arange(m - 1), indexing="ij")
    k0 = (ci * m + cj).ravel()
    # counter-clockwise in the (X0, X1) plane: (k0, k0+m, k0+m+1) and (k0, k0+m+1, k0+1)
    f1 = np.stack([k0, k0 + m, k0 + m + 1], axis=1)
    f2 = np.stack([k0, k0 + m + 1, k0 + 1], axis=1)
    faces = np.empty((2 * k0.size, 3), dtype=np.int32)
    faces[0::2] = f1
    faces[1::2] = f2
    return X, faces


def build4(n, m=None):
    """Cloth::build (Cloth.cpp:63-126) on the unit square: p00=(0,0), p01=(1,0) (JSON corner2), p10=(0,1) (corner3)."""
    m = n if m is None else m
    i, j = np.meshgrid(np.arange(n), np.arange(m), indexing="ij")
    u = i.ravel() / (n - 1.0)      # along p00 -> p10
    v = j.ravel() / (m - 1.0)      # along p00 -> p01
    # bilinear with p00=(0,0), p01=(1,0), p10=(0,1), p11=(1,1):  x = v, y = u
    grid = np.stack([v, u], axis=1)
    ci, cj = np.meshgrid(np.arange(n - 1), np.arange(m - 1), indexing="ij")
    uc = (ci.ravel() + 0.5) / (n - 1.0)
    vc = (cj.ravel() + 0.5) / (m - 1.0)
    centre = np.stack([vc, uc], axis=1)
    X = np.concatenate([grid, centre], axis=0).astype(np.float64)
    k0 = (ci * m + cj).ravel()
    kc = n * m + np.arange(k0.size)
    faces = np.empty((4 * k0.size, 3), dtype=np.int32)
    faces[0::4] = np.stack([k0, k0 + 1, kc], axis=1)
    faces[1::4] = np.stack([k0 + 1, k0 + m + 1, kc], axis=1)
    faces[2::4] = np.stack([k0 + m + 1, k0 + m, kc], axis=1)
    faces[3::4] = np.stack([k0 + m, k0, kc], axis=1)
    return X, faces


def edge_stencils(n_nodes, face_nodes):
    """(E,4) int32 stencils (n0, n1, opp(adjf0), opp(adjf1)), -1 = absent, in ArcSim mesh.edges order."""
    face_nodes = capi.i32(face_nodes).reshape(-1, 3)
    F = face_nodes.shape[0]
    out = np.empty((3 * max(F, 1), 4), dtype=np.int32)
    E = ctypes.c_int32(0)
    capi.check(capi.lib().eolc_mesh_edge_stencils(int(n_nodes), F, capi.iptr(face_nodes), ctypes.byref(E), capi.iptr(out)))
    return out[:E.value].copy()


def drape_state(X, seed=0, amp=0.05, noise=1e-3):
    """x = (X, 0.05 sin(2 pi X0) cos(2 pi X1)) + U(-1e-3, 1e-3) on all coords."""
    rng = np.random.default_rng(seed)
    x = np.zeros((X.shape[0], 3))
    x[:, :2] = X
    x[:, 2] = amp * np.sin(2 * np.pi * X[:, 0]) * np.cos(2 * np.pi * X[:, 1])
    x += rng.uniform(-noise, noise, size=x.shape)
    return x


# simulationSettingsBox.json:87-90 : whd = (1.2, 1.5, 1.0), centre (0.9175, 0.4425, -0.549), E1 = translation
BOX_WHD = np.array([1.2, 1.5, 1.0])
BOX_CENTRE = np.array([0.9175, 0.4425, -0.549])
BOX_THRESHOLD = 5e-3


def box_frame(centre=BOX_CENTRE, rot=None):
    """Column-major 4x4 frame (flattened, 16) from a centre and an optional 3x3 rotation."""
    E = np.eye(4)
    if rot is not None:
        E[:3, :3] = rot
    E[:3, 3] = centre
    return np.ascontiguousarray(E.T).ravel()   # column-major


def box_scene_state(X, seed=0, centre=BOX_CENTRE, whd=BOX_WHD, noise=1e-4):
    """the sheet lies 1e-3 inside the box top for X0 >= box xmin and dips 0.5*(xmin - X0) outside."""
    rng = np.random.default_rng(seed)
    xmin = centre[0] - 0.5 * whd[0]
    ztop = centre[2] + 0.5 * whd[2]
    x = np.zeros((X.shape[0], 3))
    x[:, :2] = X
    z = np.full(X.shape[0], ztop - 1e-3)
    out = X[:, 0] < xmin
    z[out] = ztop - 1e-3 - 0.5 * (xmin - X[out, 0])
    x[:, 2] = z
    x += rng.uniform(-noise, noise, size=x.shape)
    return x
