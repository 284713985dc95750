"""Consumer of Forces::fill on the device: b = -(M v + h f) (Cloth.cpp:345) and the collision-free CG branch
(GeneralizedSolver.cpp:120-126) against their numpy restatements in the oracle.  Tolerances: b to 1e-12 of its scale (one SpMV, FP64,
different summation order); CG solutions agree to the solver tolerance times the conditioning seen on these systems (1e-7 relative
at tol = 1e-12) and reach the requested residual."""
import numpy as np
import pytest

import eol_cloth_b200 as E

pytestmark = pytest.mark.gpu
MAT = E.Material.DEFAULT
GRAV = (0.0, 0.0, -9.8)
H = 0.5e-2


def _setup(ctx, gen, n):
    import torch
    X, fn = getattr(E.meshgen, gen)(n)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=n)
    N = X.shape[0]
    plan = E.ForcesPlan(ctx, N, fn, es, X_hint=X)
    dev = torch.device("cuda", ctx.device)
    t = dict(x=torch.from_numpy(x).to(dev), X=torch.from_numpy(X.copy()).to(dev),
             f=torch.empty(3 * N, dtype=torch.float64, device=dev), M=torch.empty(plan.nnz[0], dtype=torch.float64, device=dev),
             K=torch.empty(plan.nnz[1], dtype=torch.float64, device=dev), v=None, b=torch.empty(3 * N, dtype=torch.float64, device=dev),
             sol=torch.empty(3 * N, dtype=torch.float64, device=dev))
    torch.cuda.synchronize()
    plan.fill_dev(t["x"].data_ptr(), t["X"].data_ptr(), MAT, GRAV, H, t["f"].data_ptr(), t["M"].data_ptr(), t["K"].data_ptr())
    torch.cuda.synchronize()
    return plan, t, (X, fn, es, x), dev


@pytest.mark.parametrize("gen,n", [("regular2", 24), ("build4", 9), ("regular2", 3)])
def test_rhs_and_cg_match_oracle(ctx, oracle, gen, n):
    import torch
    plan, t, (X, fn, es, x), dev = _setup(ctx, gen, n)
    ref = oracle.forces_fill(fn, es, x, X, tuple(MAT), GRAV, H)
    dof = 3 * X.shape[0]
    v = 0.1 * np.random.default_rng(n).standard_normal(dof)
    vd = torch.from_numpy(v).to(dev)
    torch.cuda.synchronize()
    plan.rhs_dev(t["M"].data_ptr(), t["f"].data_ptr(), vd.data_ptr(), H, t["b"].data_ptr())
    torch.cuda.synchronize()
    b = t["b"].cpu().numpy()
    b_ref = oracle.cloth_rhs(ref["M"], ref["f"], v, H)
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
    it, res = plan.solve_cg_dev(t["K"].data_ptr(), t["b"].data_ptr(), t["sol"].data_ptr(), tol=1e-12, max_iter=2 * dof)
    sol = t["sol"].cpu().numpy()
    sol_ref, it_ref, res_ref = oracle.eigen_cg(ref["MDK"], b_ref, tol=1e-12)
    assert res < 1e-12 and it_ref <= it <= it_ref + 8 + max(4, it_ref // 10), (it, it_ref, res)
    assert np.abs(sol - sol_ref).max() <= 1e-7 * np.abs(sol_ref).max()
    # the solve really solves: residual of the ORACLE matrix with the GPU solution
    import scipy.sparse as sp
    o, i, vals = ref["MDK"]
    K = sp.csc_matrix((vals, i, o), shape=(dof, dof))
    assert np.linalg.norm(K @ sol + b_ref) <= 1e-9 * np.linalg.norm(b_ref)
    # bit-reproducible (fixed-order reductions)
    it2, _ = plan.solve_cg_dev(t["K"].data_ptr(), t["b"].data_ptr(), t["sol"].data_ptr(), tol=1e-12, max_iter=2 * dof)
    assert it2 == it and t["sol"].cpu().numpy().tobytes() == sol.tobytes()


def test_cg_zero_rhs_and_iteration_cap(ctx):
    import torch
    plan, t, _, dev = _setup(ctx, "regular2", 12)
    dof = t["b"].numel()
    t["b"].zero_()
    torch.cuda.synchronize()
    it, res = plan.solve_cg_dev(t["K"].data_ptr(), t["b"].data_ptr(), t["sol"].data_ptr(), tol=1e-12)
    assert it == 0 and res == 0.0 and float(t["sol"].abs().max().item()) == 0.0
    t["b"].fill_(1.0)
    torch.cuda.synchronize()
    it, res = plan.solve_cg_dev(t["K"].data_ptr(), t["b"].data_ptr(), t["sol"].data_ptr(), tol=1e-30, max_iter=16)
    assert it == 16 and res > 0.0
    with pytest.raises(E.EolcError):
        plan.solve_cg_dev(t["K"].data_ptr(), t["b"].data_ptr(), t["sol"].data_ptr(), tol=0.0)


def test_drape_steps_with_fixed_corners(ctx, oracle):
    """BASELINE configs[0]: simulationSettings.json's square cloth (Cloth::build mesh, corners 3 and 4 fixed, solver 'none', no remeshing,
    no EOL), a few collision-free steps entirely on the device — fill -> b = -(M v + h f) -> CG on the free dofs -> x += h v
    (Cloth.cpp:359-400) — against the same loop on the oracle.  The state after 5 steps agrees to 1e-9 of the cloth size."""
    import torch
    res = 5                                                  # simulationSettings.json uses 3; 5 gives interior bending stencils too
    X, fn = E.meshgen.build4(res)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    N = X.shape[0]
    x = np.c_[X, np.zeros(N)]
    x[:, 2] += 1e-3 * np.sin(3.0 * X[:, 0]) * np.cos(2.0 * X[:, 1])      # not perfectly flat: exercises bending
    # Cloth::build: grid node (i, j) -> i * res + j; corner3 = (0, 1) -> j = res - 1, corner4 = (1, 1) -> last grid node
    fixed_nodes = [res - 1, res * res - 1]
    fixed = np.zeros(3 * N, dtype=np.uint8)
    for a in fixed_nodes:
        fixed[3 * a:3 * a + 3] = 1
    plan = E.ForcesPlan(ctx, N, fn, es, X_hint=X)
    dev = torch.device("cuda", ctx.device)
    xd = torch.from_numpy(x.copy()).to(dev); Xd = torch.from_numpy(X.copy()).to(dev)
    f = torch.empty(3 * N, dtype=torch.float64, device=dev); M = torch.empty(plan.nnz[0], dtype=torch.float64, device=dev)
    K = torch.empty(plan.nnz[1], dtype=torch.float64, device=dev)
    v = torch.zeros(3 * N, dtype=torch.float64, device=dev); b = torch.empty_like(v); vn = torch.empty_like(v)
    fx = torch.from_numpy(fixed).to(dev)
    torch.cuda.synchronize()
    x_ref, v_ref = x.copy(), np.zeros(3 * N)
    for step in range(5):
        plan.fill_dev(xd.data_ptr(), Xd.data_ptr(), MAT, GRAV, H, f.data_ptr(), M.data_ptr(), K.data_ptr())
        plan.rhs_dev(M.data_ptr(), f.data_ptr(), v.data_ptr(), H, b.data_ptr())
        it, rr = plan.solve_cg_dev(K.data_ptr(), b.data_ptr(), vn.data_ptr(), tol=1e-13, fixed_ptr=fx.data_ptr())
        assert rr < 1e-13
        plan.integrate_dev(vn.data_ptr(), H, xd.data_ptr())
        v, vn = vn, v
        ref = oracle.forces_fill(fn, es, x_ref, X, tuple(MAT), GRAV, H)
        b_ref = oracle.cloth_rhs(ref["M"], ref["f"], v_ref, H)
        v_ref, _, _ = oracle.eigen_cg(ref["MDK"], b_ref, tol=1e-13, fixed=fixed.astype(bool))
        x_ref = x_ref + H * v_ref.reshape(-1, 3)
    torch.cuda.synchronize()
    got = xd.cpu().numpy()
    assert np.abs(got - x_ref).max() <= 1e-9
    assert np.all(got[fixed_nodes] == x[fixed_nodes])        # the fixed corners did not move
    assert got[:, 2].min() < x[:, 2].min() - 1e-5             # the free part falls


def test_eol_plan_rhs_cg_and_integrate(ctx, oracle):
    """The consumers on a plan with EoL nodes (scalar-CSR kernels over Eigen's arrays): b = -(M v + h f) over all 3N + 2 EoL_Count dofs;
    CG on MDK with the EoL nodes' Lagrangian dofs held (an EoL node slides along its box edge: x prescribed, X free — the unconstrained
    EOL system is singular, [I, -F] has a null space per EoL node); X += h v_X for the EoL nodes."""
    import torch
    import scipy.sparse as sp
    n = 24
    X, fn = E.meshgen.regular2(n)
    N = X.shape[0]
    es = E.meshgen.edge_stencils(N, fn)
    x = E.meshgen.drape_state(X, seed=n)
    eol = np.full(N, -1, np.int32)
    line = np.arange(1, n - 1) * n + n // 2
    eol[line] = np.random.default_rng(1).permutation(line.size)
    plan = E.ForcesPlan(ctx, N, fn, es, eol_index=eol, X_hint=X)
    dof = plan.dof
    assert dof == 3 * N + 2 * line.size
    dev = torch.device("cuda", ctx.device)
    xd, Xd = torch.from_numpy(x).to(dev), torch.from_numpy(X.copy()).to(dev)
    fd = torch.empty(dof, dtype=torch.float64, device=dev)
    Md = torch.empty(plan.nnz[0], dtype=torch.float64, device=dev)
    Kd = torch.empty(plan.nnz[1], dtype=torch.float64, device=dev)
    bd = torch.empty(dof, dtype=torch.float64, device=dev)
    sold = torch.empty(dof, dtype=torch.float64, device=dev)
    v = 0.1 * np.random.default_rng(n).standard_normal(dof)
    vd = torch.from_numpy(v).to(dev)
    torch.cuda.synchronize()
    plan.fill_dev(xd.data_ptr(), Xd.data_ptr(), MAT, GRAV, H, fd.data_ptr(), Md.data_ptr(), Kd.data_ptr())
    plan.rhs_dev(Md.data_ptr(), fd.data_ptr(), vd.data_ptr(), H, bd.data_ptr())
    torch.cuda.synchronize()
    ref = oracle.forces_fill(fn, es, x, X, tuple(MAT), GRAV, H, eol_index=eol)
    b_ref = oracle.cloth_rhs(ref["M"], ref["f"], v, H)
    b = bd.cpu().numpy()
    assert np.abs(b - b_ref).max() <= 1e-12 * np.abs(b_ref).max()
    fixed = np.zeros(dof, np.uint8)
    for a in line:
        fixed[3 * a:3 * a + 3] = 1
    fixd = torch.from_numpy(fixed).to(dev)
    torch.cuda.synchronize()
    it, res = plan.solve_cg_dev(Kd.data_ptr(), bd.data_ptr(), sold.data_ptr(), tol=1e-12, max_iter=4 * dof, fixed_ptr=fixd.data_ptr())
    sol = sold.cpu().numpy()
    sol_ref, it_ref, _ = oracle.eigen_cg(ref["MDK"], b_ref, tol=1e-12, max_iter=4 * dof, fixed=fixed.astype(bool))
    # the device counts iterations in batches of 8; rounding in the reductions may move convergence by a few iterations either way
    assert res < 1e-12 and abs(it - it_ref) <= 8 + max(8, it_ref // 5), (it, it_ref, res)
    assert np.all(sol[fixed != 0] == 0.0)
    assert np.abs(sol - sol_ref).max() <= 1e-6 * np.abs(sol_ref).max()
    o, i, vals = ref["MDK"]
    K = sp.csc_matrix((vals, i, o), shape=(dof, dof)).tocsr()
    free = np.flatnonzero(fixed == 0)
    assert np.linalg.norm((K @ sol + b_ref)[free]) <= 1e-9 * np.linalg.norm(b_ref[free])
    # x += h v (Lagrangian dofs), X += h v_X (EoL nodes): Cloth.cpp:394-407
    plan.integrate_dev(sold.data_ptr(), H, xd.data_ptr())
    plan.integrate_X_dev(sold.data_ptr(), H, Xd.data_ptr())
    torch.cuda.synchronize()
    X_new = X.copy()
    for a in line:
        X_new[a] += H * sol[3 * N + 2 * eol[a]:3 * N + 2 * eol[a] + 2]
    assert np.allclose(Xd.cpu().numpy(), X_new, rtol=0, atol=1e-15)       # the device contracts X + h v into one FMA
    assert np.allclose(xd.cpu().numpy().reshape(-1), x.reshape(-1) + H * sol[:3 * N], rtol=0, atol=1e-15)
    plan.close()
