"""The CUDA path against THE REFERENCE'S OWN Forces::fill (oracle/_ref/libforces_ref.so: Forces.cpp + UtilEOL.cpp + Compute*.cpp + ArcSim's
mesh code compiled unmodified, see tests/test_forces_ref_pin.py), directly — no restatement in between: identical compressed index
arrays, f / M / MDK within 1e-10 of the block-row scale (SURVEY §8c rule 5).  Sizes the reference finishes in seconds."""
import numpy as np
import pytest

import eol_cloth_b200 as E
from util import assert_close_tol, block_row_scale

pytestmark = pytest.mark.gpu
MAT = E.Material.DEFAULT
GRAV = (0.0, 0.0, -9.8)
H = 0.5e-2


def _gpu_fill(ctx, X, fn, x, eol=None):
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    plan = E.ForcesPlan(ctx, X.shape[0], fn, es, eol_index=eol, X_hint=X)
    try:
        f, Mv, Kv = plan.fill(x, X, MAT, GRAV, H)
        return dict(dof=plan.dof, f=f, M=(*plan.pattern(0), Mv), MDK=(*plan.pattern(1), Kv))
    finally:
        plan.close()


def _compare(ref, got, n_nodes, what):
    assert ref["dof"] == got["dof"], what
    for k in ("M", "MDK"):
        assert np.array_equal(ref[k][0], got[k][0]) and np.array_equal(ref[k][1], got[k][1]), f"{what}: {k} index arrays differ"
        assert_close_tol(got[k][2], ref[k][2], block_row_scale(ref[k][0], ref[k][2], n_nodes), 1e-10, f"{what} {k}")
    assert_close_tol(got["f"], ref["f"], np.abs(ref["f"]).max(), 1e-10, f"{what} f")


@pytest.mark.parametrize("gen,n", [("regular2", 3), ("regular2", 64), ("build4", 7), ("build4", 40), ("regular2", 129), ("regular2", 256)])
def test_gpu_fill_equals_the_reference_fill(ctx, oracle, gen, n):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.drape_state(X, seed=n)
    _compare(oracle.ref_forces_fill(fn, x, X, tuple(MAT), GRAV, H), _gpu_fill(ctx, X, fn, x), X.shape[0], f"{gen}{n}")


@pytest.mark.parametrize("n,kind", [(12, "line"), (24, "scattered"), (6, "all")])
def test_gpu_eol_fill_equals_the_reference_fill(ctx, oracle, n, kind):
    X, fn = E.meshgen.regular2(n)
    N = X.shape[0]
    eol = np.full(N, -1, np.int32)
    if kind == "line":
        sel = np.arange(1, n - 1) * n + n // 2
        eol[sel] = np.arange(sel.size)
    elif kind == "scattered":
        sel = np.random.default_rng(4).permutation(N)[:40]
        eol[sel] = np.random.default_rng(5).permutation(40)
    else:
        eol[:] = np.arange(N)
    x = E.meshgen.drape_state(X, seed=3)
    ref = oracle.ref_forces_fill(fn, x, X, tuple(MAT), GRAV, H, eol_index=eol)
    assert ref["dof"] == 3 * N + 2 * int((eol >= 0).sum())
    _compare(ref, _gpu_fill(ctx, X, fn, x, eol), N, f"EOL {kind}")


def test_gpu_fill_512_equals_the_reference_fill(ctx, oracle):
    """The largest size at which the reference's own code is run in the suite (1.3 M elements, 0.2 G triplets, ~20 s on one thread;
    the whole 1024^2 sheet goes through the pinned restatement, tests/test_forces_gpu.py::test_fullsize_1024_values_match_oracle)."""
    X, fn = E.meshgen.regular2(512)
    x = E.meshgen.drape_state(X, seed=0)
    ref = oracle.ref_forces_fill(fn, x, X, tuple(MAT), GRAV, H)
    assert len(ref["MDK"][2]) == 30560364 and len(ref["M"][2]) == 16478226
    _compare(ref, _gpu_fill(ctx, X, fn, x), X.shape[0], "regular2 512")
