"""Pins the collision oracle (oracle/cd_ref.cpp, the restatement the CUDA path is compared with) to the REFERENCE'S OWN code:
oracle/_ref/libbtc_ref.so is /root/reference/src/{boxTriCollision,Collisions,raytri}.cpp compiled UNMODIFIED against
oracle/mini_eigen (oracle/Makefile).  Every field of every contact must be identical BIT FOR BIT, list order included.

The reference's `int` edge hash (boxTriCollision.cpp:167-169) overflows for (3F+1) N >= 2^31 (N > ~19 k nodes), so the
reference itself is only defined on the small cases and the 64x64 ensemble scenes — all of which run here; above that size
the oracle's 64-bit key continues the same order (tests/test_cd_gpu.py).

The second library, libbtc_ref_scalar.so, is the same reference code with the OTHER candidate for Eigen's 3-term reduction
order (p0+(p1+p2), a non-vectorised Eigen build).  The contact SET — count, order and every integer field — is the same under
both; only last bits of real fields move.  So "bit-exact contact index sets" does not hinge on that detail of Eigen."""
import numpy as np
import pytest

import eol_cloth_b200 as E
from eol_cloth_b200.collisions import make_obstacles
from util import INT_FIELDS, REAL_FIELDS, assert_contacts_equal

THR = E.meshgen.BOX_THRESHOLD
C3B = np.array([0.9175, -0.25, -0.549])  # SURVEY §8d variant 3b: a box corner under the cloth


def _rot(axis, ang):
    axis = np.asarray(axis, float) / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K


def _assert_bits(a, b, what):
    assert len(a) == len(b), f"{what}: {len(a)} contacts (reference) vs {len(b)} (oracle)"
    for f in a.dtype.names:
        assert a[f].tobytes() == b[f].tobytes(), f"{what}: field {f} differs in {int(np.count_nonzero(a[f] != b[f]))} places"


def _scene(gen, n, centre, rot, seed):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
    R = None if rot is None else _rot(*rot)
    return X, fn, x, E.meshgen.BOX_WHD[None], E.meshgen.box_frame(np.asarray(centre), R)[None]


# the cases of tests/test_cd_gpu.py + general rotations + the ensemble's scene size
CASES = [
    ("regular2", 24, tuple(E.meshgen.BOX_CENTRE), None, 0),
    ("build4", 16, tuple(C3B), None, 1),
    ("regular2", 40, tuple(C3B), None, 2),
    ("regular2", 33, (0.7, 0.3, -0.549), ((0, 0, 1), 0.3), 3),
    ("build4", 21, (0.5, 0.5, -0.549), ((1, 2, 0.5), 0.05), 4),
    ("regular2", 31, (0.6, 0.4, -0.53), ((1, -1, 0.3), 0.11), 5),
    ("regular2", 64, tuple(C3B), None, 7),
    ("regular2", 64, tuple(E.meshgen.BOX_CENTRE), None, 8),
    ("build4", 48, (0.7, 0.3, -0.549), ((0.2, 0.1, 1), 0.7), 9),
    ("regular2", 3, tuple(C3B), None, 10),      # N = 9 > F = 8 (the faceNors2.col(vertex id) quirk at its limit)
    ("regular2", 2, tuple(C3B), None, 11),
]


@pytest.mark.parametrize("gen,n,centre,rot,seed", CASES)
def test_oracle_equals_reference_box_scenes(oracle, gen, n, centre, rot, seed):
    X, fn, x, whd, Em = _scene(gen, n, centre, rot, seed)
    if n <= 3:  # keep every vertex with index >= F out of the box: the reference would read faceNors2 out of range (UB)
        x[len(fn):, 2] += 1.0
    for which, flag, remap in ((1, 1, 1), (0, 0, 0)):
        ref = oracle.ref_cd(fn, x, THR, None, None, whd, Em, which)
        got = oracle.cd(fn, x, THR, None, None, whd, Em, flag, remap)
        _assert_bits(ref, got, f"{gen}{n} {'CD' if which else 'CD2'}")
    if n >= 16:
        assert len(ref) > 0


def test_oracle_equals_reference_points_and_two_boxes(oracle):
    X, fn = E.meshgen.regular2(30)
    x = E.meshgen.box_scene_state(X, seed=5, centre=C3B)
    pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3, x[100] - 2e-3])
    pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0], [0, 0.6, 0.8]])
    whd = np.stack([E.meshgen.BOX_WHD, [0.3, 0.3, 0.3]])
    Em = np.stack([E.meshgen.box_frame(C3B), E.meshgen.box_frame(np.array([0.15, 0.8, -0.36]), _rot((0.3, 1, 0.2), 0.4))])
    eol = (np.arange(len(x)) % 7 == 0).astype(np.int32)  # Node::EoL flags: copied by CD, never read by the narrow phase
    for which, flag, remap in ((1, 1, 1), (0, 0, 0)):
        ref = oracle.ref_cd(fn, x, THR, pxyz, pn, whd, Em, which, eol=eol)
        got = oracle.cd(fn, x, THR, pxyz, pn, whd, Em, flag, remap)
        _assert_bits(ref, got, f"points+2boxes which={which}")
    kinds = set(zip(ref["count1"].tolist(), ref["count2"].tolist()))
    assert {(1, 3), (2, 2), (3, 1)} <= kinds
    cd = oracle.ref_cd(fn, x, THR, pxyz, pn, whd, Em, 1)
    # CD's index remap (Collisions.cpp:39-48) happened in the reference's own code: second box's corners land at 4 + 20 + i
    b2 = cd[(cd["count1"] == 1) & (cd["n_edge1"] == 3)]
    assert len(b2) and (b2["verts1"][:, 0] >= 4).all()


def test_point_sections_alone(oracle):
    """pointTriCollision with and without its EOL branch (CD passes true, CD2 false: Collisions.cpp:30, :72)."""
    X, fn = E.meshgen.build4(9)
    x = E.meshgen.drape_state(X, seed=3)
    rng = np.random.default_rng(4)
    pick = rng.choice(len(x), 12, replace=False)
    pxyz = np.r_[x[pick] + rng.uniform(-3e-3, 3e-3, (12, 3)), x[fn[::7]].mean(1) - [0, 0, 2e-3], [[5.0, 5.0, 5.0]]]
    pn = np.tile([0.0, 0.0, 1.0], (len(pxyz), 1))
    pn[::3] = [0.0, 0.6, -0.8]
    for eolflag in (True, False):
        ref = oracle.ref_btc_points(fn, x, THR, pxyz, pn, eolflag)
        got = oracle.cd(fn, x, THR, pxyz, pn, None, None, int(eolflag), 0)
        _assert_bits(ref, got, f"points EOL={eolflag}")
        assert len(ref) > 0
    assert (oracle.ref_btc_points(fn, x, THR, pxyz, pn, True)["count1"] == 3).any()


@pytest.mark.parametrize("gen,n", [("regular2", 17), ("build4", 11), ("regular2", 64)])
def test_edge_table_equals_reference(oracle, gen, n):
    """btc::createEdges: order, verts, faces, normals (from the UNPERTURBED vertices) — the oracle's 64-bit key gives the reference's
    order wherever the reference's int key is defined."""
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.drape_state(X, seed=n)
    tab_r, nrm_r, internal, angle = oracle.ref_btc_edges(fn, x)
    tab_o, nrm_o = oracle.cd_edges(fn, x)
    assert np.array_equal(tab_r, tab_o)
    assert nrm_r.tobytes() == nrm_o.tobytes()
    assert np.array_equal(internal == 1, tab_r[:, 3] >= 0)
    assert (angle[internal == 0] == 1e9).all()


@pytest.mark.parametrize("rot", [None, ((0, 0, 1), 0.3), ((1, 2, 0.5), 0.05), ((-0.3, 0.9, 0.4), 2.1)])
def test_box_tables_equal_reference(oracle, rot):
    """createBox: E1*S, E*verts1_ (4-term products), face normals, angle-weighted vertex normals (acos), edge angles."""
    R = None if rot is None else _rot(*rot)
    Em = E.meshgen.box_frame(np.array([0.31, -0.27, 0.113]), R)
    whd = np.array([1.2, 1.5, 0.7])
    for a, b in zip(oracle.ref_btc_boxtables(whd, Em), oracle.box(whd, Em)):
        assert a.tobytes() == b.tobytes()


def test_ensemble_scenes_equal_reference(oracle):
    """BASELINE configs[4]: 64x64 scenes (state seed = scene id) against the same box — sampled scenes, CD2."""
    X, fn = E.meshgen.regular2(64)
    whd, Em = E.meshgen.BOX_WHD[None], E.meshgen.box_frame(C3B)[None]
    for scene in (0, 1, 77, 1023, 4095):
        x = E.meshgen.box_scene_state(X, seed=scene, centre=C3B)
        _assert_bits(oracle.ref_cd(fn, x, THR, None, None, whd, Em, 0), oracle.cd(fn, x, THR, None, None, whd, Em, 0, 0), f"scene {scene}")


@pytest.mark.parametrize("gen,n,centre,rot,seed", CASES[:9])
def test_contact_set_independent_of_eigen_reduction_order(oracle, gen, n, centre, rot, seed):
    """Same reference sources, Eigen's other reduction order: identical contact set; real fields move by rounding only."""
    X, fn, x, whd, Em = _scene(gen, n, centre, rot, seed)
    a = oracle.ref_cd(fn, x, THR, None, None, whd, Em, 1)
    b = oracle.ref_cd(fn, x, THR, None, None, whd, Em, 1, scalar_redux=True)
    assert_contacts_equal(b, a, real_tol=1e-12, what=f"{gen}{n} scalar-order vs vector-order reference")
    for f in INT_FIELDS:
        assert np.array_equal(a[f], b[f])


def test_reference_hash_limit_is_reported(oracle):
    assert oracle.ref_hash_defined(64 * 64, 2 * 63 * 63)
    assert not oracle.ref_hash_defined(256 * 256, 2 * 255 * 255)
    X, fn = E.meshgen.regular2(160)
    with pytest.raises(ValueError):
        oracle.ref_btc_edges(fn, np.c_[X, np.zeros(len(X))])


def test_golden_cd_fixtures_come_from_the_reference(oracle):
    """tests/golden/cd_*.npz are written from libbtc_ref.so (tests/golden/make_golden.py) — check they still are its output."""
    import os
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    for name, gen, n, centre, seed, points in (("cd_regular2_n24", "regular2", 24, E.meshgen.BOX_CENTRE, 0, False),
                                               ("cd_build4_n16_corner", "build4", 16, C3B, 1, True)):
        g = np.load(os.path.join(gold, name + ".npz"))
        X, fn = getattr(E.meshgen, gen)(n)
        x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
        pxyz = pn = None
        if points:
            pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3])
            pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]])
        for key, which in (("cd", 1), ("cd2", 0)):
            ref = oracle.ref_cd(fn, x, THR, pxyz, pn, E.meshgen.BOX_WHD[None], E.meshgen.box_frame(np.asarray(centre))[None], which)
            _assert_bits(ref, g[key], f"{name} {key}")
        assert str(g["source"]) == "libbtc_ref"
