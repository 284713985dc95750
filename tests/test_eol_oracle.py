"""EOL branch of Forces::fill: the oracle's line-by-line restatement of fillEOLInertia /
fillEOLMembrane / fillEOLBending and the EOL scatters (oracle/forces_ref.cpp) against an independent dense derivation.

The reference's expansions are the congruence  K_exp = G^T K G,  f_exp = G^T f  with, per vertex v of the element,
G_v = [I_3, -F] (3 x 5) when v is an EoL node and [I_3] otherwise (Forces.cpp:177-329, 580-683: the blocks it writes are
Kxx, -F^T Kxx, -Kxx F and F^T Kxx F).  The dense side below builds exactly that with numpy from the reference's own
Compute* object code, so the two derivations only share the element kernels."""
import numpy as np
import pytest

import eol_cloth_b200 as E

MAT = (0.05, 50.0, 0.01, 1.0e-5, 0.0, 1.0)
GRAV = (0.0, 0.0, -9.8)
H = 0.5e-2


def eol_case(n=5, eol_nodes=(12, 6, 18, 7), seed=4, gen="regular2"):
    X, fn = getattr(E.meshgen, gen)(n)
    N = X.shape[0]
    es = E.meshgen.edge_stencils(N, fn)
    x = E.meshgen.drape_state(X, seed=seed)
    eol = np.full(N, -1, np.int32)
    for k, a in enumerate(eol_nodes):
        eol[a] = k
    return X, fn, es, x, eol


def dense_reference(oracle, X, fn, es, x, eol, mat=MAT, grav=GRAV, h=H):
    N = X.shape[0]
    n_eol = int(eol.max()) + 1
    dof = 3 * N + 2 * n_eol
    rho, e, nu, beta, _, dB = mat
    dhh = dB * h * h
    f = np.zeros(dof); M = np.zeros((dof, dof)); K = np.zeros((dof, dof))
    maskM = np.zeros((dof, dof), bool); maskK = np.zeros((dof, dof), bool)

    def defgrad(tri):
        a, b, c = tri
        Dx = np.c_[x[b] - x[a], x[c] - x[a]]
        DX = np.c_[X[b] - X[a], X[c] - X[a]]
        return Dx @ np.linalg.inv(DX)

    def G_and_dofs(nodes, F):
        cols = []; blocks = []
        for v, a in enumerate(nodes):
            cols += [3 * a, 3 * a + 1, 3 * a + 2]
            if eol[a] >= 0:
                cols += [3 * N + 2 * eol[a], 3 * N + 2 * eol[a] + 1]
        G = np.zeros((3 * len(nodes), len(cols)))
        c = 0
        for v, a in enumerate(nodes):
            G[3 * v:3 * v + 3, c:c + 3] = np.eye(3); c += 3
            if eol[a] >= 0:
                G[3 * v:3 * v + 3, c:c + 2] = -F; c += 2
        return G, np.array(cols)

    for tri in fn:
        xs = [x[a] for a in tri]; Xs = [X[a] for a in tri]
        P, Q = oracle.face_frame(*xs, *Xs)
        _, fm, Km = oracle.compute_membrane(*xs, *Xs, e, nu, P, Q)
        _, fi, Mi = oracle.compute_inertial(*xs, *Xs, grav, rho)
        G, d = G_and_dofs(tri, defgrad(tri))
        f[d] += G.T @ (fm + fi)
        M[np.ix_(d, d)] += G.T @ Mi @ G
        K[np.ix_(d, d)] += G.T @ (Mi + dhh * Km) @ G
        maskM[np.ix_(d, d)] = True; maskK[np.ix_(d, d)] = True
    faces_of = {tuple(sorted(t)): t for t in fn.tolist()}
    for s in es:
        if s[2] < 0 or s[3] < 0:
            continue
        xs = [x[a] for a in s]; Xs = [X[a] for a in s]
        _, fb, Kb = oracle.compute_bending(*xs, *Xs, beta)
        any_eol = any(eol[a] >= 0 for a in s)
        F = None
        if any_eol:
            F = 0.5 * (defgrad(faces_of[tuple(sorted((s[0], s[1], s[2])))]) + defgrad(faces_of[tuple(sorted((s[0], s[1], s[3])))]))
        G, d = G_and_dofs(s, F)
        if any_eol:
            f[d] += G.T @ fb            # the Lagrangian branch drops the bending force (Forces.cpp:885-908)
        K[np.ix_(d, d)] += G.T @ (dhh * Kb) @ G
        maskK[np.ix_(d, d)] = True
    return dof, f, M, K, maskM, maskK


def to_dense(mat, dof):
    o, i, v = mat
    D = np.zeros((dof, dof)); Pm = np.zeros((dof, dof), bool)
    for c in range(dof):
        D[i[o[c]:o[c + 1]], c] = v[o[c]:o[c + 1]]
        Pm[i[o[c]:o[c + 1]], c] = True
        assert np.all(np.diff(i[o[c]:o[c + 1]]) > 0)
    return D, Pm


@pytest.mark.parametrize("gen,n,eol_nodes", [("regular2", 5, (12, 6, 18, 7)), ("regular2", 4, (5,)), ("build4", 3, (4, 9, 10, 1)),
                                             ("regular2", 3, tuple(range(9)))])
def test_eol_oracle_is_the_congruence(oracle, gen, n, eol_nodes):
    X, fn, es, x, eol = eol_case(n, eol_nodes, gen=gen)
    ref = oracle.forces_fill(fn, es, x, X, MAT, GRAV, H, eol_index=eol)
    dof, f, M, K, maskM, maskK = dense_reference(oracle, X, fn, es, x, eol)
    assert ref["dof"] == dof == 3 * X.shape[0] + 2 * len(eol_nodes)
    assert np.abs(ref["f"] - f).max() <= 1e-13 * np.abs(f).max()
    for name, D, mask in (("M", M, maskM), ("MDK", K, maskK)):
        got, pat = to_dense(ref[name], dof)
        assert np.array_equal(pat, mask), name + " pattern"
        # mirrored triplets: exactly symmetric except inside the (X_P, X_P) diagonal 2x2 blocks, which F^T K F fills entry by entry
        asym = np.argwhere(got != got.T)
        N3 = 3 * X.shape[0]
        assert all(r >= N3 and c >= N3 and (r - N3) // 2 == (c - N3) // 2 for r, c in asym), name
        assert np.abs(got - got.T).max() <= 1e-15 * np.abs(got).max()
        assert np.abs(got - D).max() <= 1e-13 * np.abs(D).max(), name
        assert not np.isnan(got).any()


def test_eol_oracle_without_eol_nodes_is_the_lagrangian_branch(oracle):
    X, fn, es, x, eol = eol_case(6, ())
    a = oracle.forces_fill(fn, es, x, X, MAT, GRAV, H)
    b = oracle.forces_fill(fn, es, x, X, MAT, GRAV, H, eol_index=eol)
    assert a["f"].tobytes() == b["f"].tobytes()
    for name in ("M", "MDK"):
        for u, v in zip(a[name], b[name]):
            assert u.tobytes() == v.tobytes()


def test_eol_oracle_lagrangian_part_is_untouched_except_bending_force(oracle):
    """x-x blocks are the Lagrangian ones bit for bit (Forces.cpp:414-429 push the same values in the same order); f differs only by the
    bending force of the stencils that hold an EoL node (:750-760)."""
    X, fn, es, x, eol = eol_case(6, (14, 15, 21))
    N = X.shape[0]
    a = oracle.forces_fill(fn, es, x, X, MAT, GRAV, H)
    b = oracle.forces_fill(fn, es, x, X, MAT, GRAV, H, eol_index=eol)
    for name in ("M", "MDK"):
        A, _ = to_dense(a[name], 3 * N)
        B, _ = to_dense(b[name], b["dof"])
        assert np.array_equal(A, B[:3 * N, :3 * N]), name
    touched = np.zeros(N, bool)
    for s in es:
        if s[2] >= 0 and s[3] >= 0 and (eol[s] >= 0).any():
            touched[s] = True
    d = (a["f"] != b["f"][:3 * N]).reshape(N, 3).any(axis=1)
    assert d.any() and not (d & ~touched).any()
