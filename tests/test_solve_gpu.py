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
