"""Per-step derived mesh data: world-space face / node normals of compute_ws_data
(/root/reference/src/external/ArcSim/mesh.cpp:135-143, geometry.cpp:302-316).  CPU: the numpy oracle against a literal loop over the
pointer-graph formulation; GPU: eolc_mesh_normals[_dev] against the oracle."""
import math

import numpy as np
import pytest

import eol_cloth_b200 as E


def _literal(fn, x):
    """geometry.cpp:302-316 / mesh.cpp:135-140 written as the reference's loops (scalar Python, small meshes only)."""
    def sub(a, b): return [a[0] - b[0], a[1] - b[1], a[2] - b[2]]
    def dot(a, b):
        d = 0.0
        for i in range(3):
            d += a[i] * b[i]
        return d
    def cross(u, v): return [u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]]
    def div(u, a):          # vectors.hpp:104: u / a is u * (1 / a) — one reciprocal, then products (pinned to the reference's own
        r = 1 / a           # compiled code by tests/test_forces_ref_pin.py::test_normals_of_compute_ws_data)
        return [r * u[0], r * u[1], r * u[2]]
    def normalize(u):
        m = math.sqrt(dot(u, u))
        return [0.0, 0.0, 0.0] if m == 0 else div(u, m)
    x = x.tolist()
    face_n = [normalize(cross(sub(x[f[1]], x[f[0]]), sub(x[f[2]], x[f[0]]))) for f in fn.tolist()]
    adjf = [[] for _ in x]
    for i, f in enumerate(fn.tolist()):        # Mesh::add(Face): include(face, v->adjf), mesh.cpp:372
        for v in f:
            adjf[v].append(i)
    node_n = []
    for a in range(len(x)):
        n = [0.0, 0.0, 0.0]
        for i in adjf[a]:
            f = fn[i].tolist()
            j = f.index(a); j1 = (j + 1) % 3; j2 = (j + 2) % 3
            e1 = sub(x[f[j1]], x[a]); e2 = sub(x[f[j2]], x[a])
            c = div(cross(e1, e2), 2 * dot(e1, e1) * dot(e2, e2))
            n = [n[0] + c[0], n[1] + c[1], n[2] + c[2]]
        node_n.append(normalize(n))
    return np.array(face_n).reshape(-1, 3), np.array(node_n).reshape(-1, 3)


def _case(gen, n, seed=0, isolated=False):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.drape_state(X, seed=seed)
    if isolated:        # one node no face refers to: normal 0 (normalize of the zero vector, vectors.hpp:111)
        x = np.r_[x, [[0.3, 0.3, 0.3]]]
    return fn, x


@pytest.mark.parametrize("gen,n,isolated", [("regular2", 2, False), ("regular2", 7, True), ("build4", 5, False)])
def test_oracle_normals_match_the_literal_loops(oracle, gen, n, isolated):
    fn, x = _case(gen, n, seed=n, isolated=isolated)
    fa, na = oracle.mesh_normals(fn, x)
    fb, nb = _literal(fn, x)
    assert fa.tobytes() == fb.tobytes() and na.tobytes() == nb.tobytes()
    assert np.allclose(np.linalg.norm(fa, axis=1), 1.0, atol=1e-15)
    if isolated:
        assert np.array_equal(na[-1], np.zeros(3))


def test_oracle_normals_flat_sheet(oracle):
    """A flat sheet: normals (0, 0, 1) up to ONE rounding — ArcSim normalises with u * (1 / m) (vectors.hpp:104), and (1 / m) * m need
    not be exactly 1 (n = 6: m = 0.04, z = 1 + 2^-52); the reference's own compiled code gives the same bits."""
    for n in (6, 9):
        X, fn = E.meshgen.regular2(n)
        x = np.c_[X, np.zeros(len(X))]
        fa, na = oracle.mesh_normals(fn, x)
        assert np.all(fa[:, :2] == 0) and np.all(na[:, :2] == 0)
        assert np.abs(fa[:, 2] - 1).max() <= 2.3e-16 and np.abs(na[:, 2] - 1).max() <= 2.3e-16
        _, fr, nr = oracle.ref_mesh_data(fn, x, X)
        assert fa.tobytes() == fr.tobytes() and na.tobytes() == nr.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("gen,n,isolated", [("regular2", 2, False), ("regular2", 7, True), ("build4", 9, False), ("regular2", 256, False)])
def test_gpu_normals_match_oracle(ctx, oracle, gen, n, isolated):
    """The kernels use explicit round-to-nearest operations in the reference's order (no FMA contraction), sqrt and division are IEEE:
    the normals are compared bit for bit."""
    fn, x = _case(gen, n, seed=n, isolated=isolated)
    es = E.meshgen.edge_stencils(x.shape[0], fn)
    plan = E.ForcesPlan(ctx, x.shape[0], fn, es)
    fa, na = oracle.mesh_normals(fn, x)
    fg, ng = plan.normals(x)
    assert fg.tobytes() == fa.tobytes(), np.abs(fg - fa).max()
    assert ng.tobytes() == na.tobytes(), np.abs(ng - na).max()
    plan.close()


@pytest.mark.gpu
def test_gpu_normals_dev_after_integrate(ctx, oracle):
    """Device-resident use: x += h v on the device (Cloth.cpp:394-400), then the normals of the new state without a host round trip."""
    import torch
    fn, x = _case("regular2", 64, seed=5)
    N = x.shape[0]
    es = E.meshgen.edge_stencils(N, fn)
    plan = E.ForcesPlan(ctx, N, fn, es)
    dev = torch.device("cuda", ctx.device)
    v = np.random.default_rng(2).standard_normal((N, 3))
    xd = torch.from_numpy(x).to(dev); vd = torch.from_numpy(v).to(dev)
    fnd = torch.empty((len(fn), 3), dtype=torch.float64, device=dev); nnd = torch.empty((N, 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    plan.integrate_dev(vd.data_ptr(), 1e-2, xd.data_ptr())
    plan.normals_dev(xd.data_ptr(), fnd.data_ptr(), nnd.data_ptr())
    plan.normals_dev(xd.data_ptr(), None, nnd.data_ptr())
    torch.cuda.synchronize()
    xh = xd.cpu().numpy()
    assert np.allclose(xh, x + 1e-2 * v, rtol=0, atol=1e-15)
    fa, na = oracle.mesh_normals(fn, xh)
    assert fnd.cpu().numpy().tobytes() == fa.tobytes() and nnd.cpu().numpy().tobytes() == na.tobytes()
    plan.close()
