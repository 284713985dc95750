"""The hand-derived element arithmetic the CUDA kernels use (csrc/elements.cuh), compiled for the host, against
the reference's generated code (oracle).  Tolerance: 1e-12 of the block scale (the GPU budget is 1e-10)."""
import ctypes

import numpy as np

dp = ctypes.POINTER(ctypes.c_double)


def d(a):
    return a.ctypes.data_as(dp)


EDGE_PAIRS = [(0, 0), (1, 1), (2, 2), (3, 3), (0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
FACE_PAIRS = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]


def test_bending_blocks(oracle, hostmath):
    hostmath.hostmath_edge.argtypes = [dp] * 8 + [ctypes.c_double] * 2 + [dp]
    rng = np.random.default_rng(0)
    worst = 0.0
    for trial in range(300):
        X = np.array([[0, 0], [1, 0], [0.3, 0.8], [0.6, -0.9]]) + 0.1 * rng.standard_normal((4, 2))
        amp = (0.3, 1e-3, 1e-6)[trial % 3]
        x = np.ascontiguousarray(np.c_[X, np.zeros(4)] + amp * rng.standard_normal((4, 3)))
        X = np.ascontiguousarray(X)
        K = np.zeros(90)
        dhh = 2.5e-5
        hostmath.hostmath_edge(d(x[0]), d(x[1]), d(x[2]), d(x[3]), d(X[0]), d(X[1]), d(X[2]), d(X[3]), 1e-5, dhh, d(K))
        _, _, Kr = oracle.compute_bending(*x, *X, 1e-5)
        Kr = Kr * dhh
        sc = np.abs(Kr).max()
        for b, (i, j) in enumerate(EDGE_PAIRS):
            worst = max(worst, np.abs(K[9 * b:9 * b + 9].reshape(3, 3) - Kr[3 * i:3 * i + 3, 3 * j:3 * j + 3]).max() / sc)
    assert worst < 1e-12, worst


def test_face_blocks(oracle, hostmath):
    hostmath.hostmath_face.argtypes = [dp] * 6 + [ctypes.c_double] * 3 + [dp, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(1)
    g = np.array([0, 0, -9.8])
    worstK = worstf = 0.0
    for trial in range(300):
        X = np.array([[0, 0], [1, 0], [0.3, 0.8]]) + 0.1 * rng.standard_normal((3, 2))
        if trial % 5 == 0:
            X = X[[0, 2, 1]]            # negative rest orientation
        amp = (0.3, 1e-3)[trial % 2]
        x = np.ascontiguousarray(np.c_[X, np.zeros(3)] + amp * rng.standard_normal((3, 3)))
        X = np.ascontiguousarray(X)
        f9, t8, K = np.zeros(9), np.zeros(1), np.zeros(54)
        dhh = 2.5e-5
        hostmath.hostmath_face(d(x[0]), d(x[1]), d(x[2]), d(X[0]), d(X[1]), d(X[2]), 50.0, 0.01, 0.05, d(g), dhh, d(f9), d(t8), d(K))
        P, Q = oracle.face_frame(*x, *X)
        _, fm, Km = oracle.compute_membrane(*x, *X, 50.0, 0.01, P, Q)
        _, fi, Mi = oracle.compute_inertial(*x, *X, g, 0.05)
        ref = Mi + dhh * Km
        sc = np.abs(ref).max()
        for b, (i, j) in enumerate(FACE_PAIRS):
            worstK = max(worstK, np.abs(K[9 * b:9 * b + 9].reshape(3, 3) - ref[3 * i:3 * i + 3, 3 * j:3 * j + 3]).max() / sc)
        worstf = max(worstf, np.abs(f9 - (fm + fi)).max() / max(np.abs(fm).max(), np.abs(fi).max()))
        assert t8[0] / 12 == Mi[0, 0] and t8[0] / 24 == Mi[0, 3]
    assert worstK < 1e-12 and worstf < 1e-11, (worstK, worstf)


def test_row_forms(oracle, hostmath):
    """Row forms used by the owner-computes assembly kernel: every block row of the element matrices."""
    hostmath.hostmath_edge_row.argtypes = [ctypes.c_int] + [dp] * 8 + [ctypes.c_double] * 2 + [dp]
    hostmath.hostmath_face_row.argtypes = [ctypes.c_int] + [dp] * 6 + [ctypes.c_double] * 3 + [dp, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(2)
    g = np.array([0.3, -0.2, -9.8])
    dhh = 2.5e-5
    worstE = worstF = worstf = 0.0
    for trial in range(200):
        X = np.array([[0, 0], [1, 0], [0.3, 0.8], [0.6, -0.9]]) + 0.1 * rng.standard_normal((4, 2))
        amp = (0.3, 1e-3, 1e-6)[trial % 3]
        x = np.ascontiguousarray(np.c_[X, np.zeros(4)] + amp * rng.standard_normal((4, 3)))
        X = np.ascontiguousarray(X)
        _, _, Kr = oracle.compute_bending(*x, *X, 1e-5)
        Kr = Kr * dhh
        sc = np.abs(Kr).max()
        for i in range(4):
            K = np.zeros(36)
            hostmath.hostmath_edge_row(i, d(x[0]), d(x[1]), d(x[2]), d(x[3]), d(X[0]), d(X[1]), d(X[2]), d(X[3]), 1e-5, dhh, d(K))
            worstE = max(worstE, np.abs(K.reshape(4, 3, 3).transpose(1, 0, 2).reshape(3, 12) - Kr[3 * i:3 * i + 3]).max() / sc)
        P, Q = oracle.face_frame(*x[:3], *X[:3])
        _, fm, Km = oracle.compute_membrane(*x[:3], *X[:3], 50.0, 0.01, P, Q)
        _, fi, Mi = oracle.compute_inertial(*x[:3], *X[:3], g, 0.05)
        ref = Mi + dhh * Km
        rows = []
        for v in range(3):
            f3, t8, K = np.zeros(3), np.zeros(1), np.zeros(27)
            hostmath.hostmath_face_row(v, d(x[0]), d(x[1]), d(x[2]), d(X[0]), d(X[1]), d(X[2]), 50.0, 0.01, 0.05, d(g), dhh, d(f3), d(t8), d(K))
            R = K.reshape(3, 3, 3).transpose(1, 0, 2).reshape(3, 9)
            rows.append(R)
            worstF = max(worstF, np.abs(R - ref[3 * v:3 * v + 3]).max() / np.abs(ref).max())
            worstf = max(worstf, np.abs(f3 - (fm + fi)[3 * v:3 * v + 3]).max() / max(np.abs(fm).max(), np.abs(fi).max()))
            assert abs(t8[0] / 12 - Mi[0, 0]) <= 1e-15 * abs(Mi[0, 0])
        full = np.vstack(rows)
        assert np.array_equal(full, full.T)          # face rows are exact transposes of each other
    assert worstE < 1e-12 and worstF < 1e-12 and worstf < 1e-11, (worstE, worstF, worstf)


def test_tile_forms(oracle, hostmath):
    """Element forms of the tiles pipeline: reduced-coordinate bending (6 computed + 4 derived blocks) and the triangle."""
    hostmath.hostmath_edge_tile.argtypes = [dp] * 8 + [ctypes.c_double] * 2 + [dp]
    hostmath.hostmath_face_tile.argtypes = [dp] * 6 + [ctypes.c_double] * 3 + [dp, ctypes.c_double, dp, dp, dp]
    rng = np.random.default_rng(3)
    g = np.array([0.3, -0.2, -9.8])
    dhh = 2.5e-5
    worstE = worstF = worstf = 0.0
    for trial in range(400):
        X = np.array([[0, 0], [1, 0], [0.3, 0.8], [0.6, -0.9]]) + 0.1 * rng.standard_normal((4, 2))
        amp = (0.3, 1e-3, 1e-6, 0.0)[trial % 4]            # 0.0: perfectly flat sheet (D = 1)
        x = np.ascontiguousarray(np.c_[X, np.zeros(4)] + amp * rng.standard_normal((4, 3)))
        if trial % 7 == 0:
            x = x * 1e-2 + 3.0                              # small elements far from the origin
            X = X * 1e-2
        X = np.ascontiguousarray(X)
        K = np.zeros(90)
        hostmath.hostmath_edge_tile(d(x[0]), d(x[1]), d(x[2]), d(x[3]), d(X[0]), d(X[1]), d(X[2]), d(X[3]), 1e-5, dhh, d(K))
        _, _, Kr = oracle.compute_bending(*x, *X, 1e-5)
        Kr = Kr * dhh
        sc = np.abs(Kr).max()
        for b, (i, j) in enumerate(EDGE_PAIRS):
            worstE = max(worstE, np.abs(K[9 * b:9 * b + 9].reshape(3, 3) - Kr[3 * i:3 * i + 3, 3 * j:3 * j + 3]).max() / sc)
        f9, t8, Kf = np.zeros(9), np.zeros(1), np.zeros(54)
        hostmath.hostmath_face_tile(d(x[0]), d(x[1]), d(x[2]), d(X[0]), d(X[1]), d(X[2]), 50.0, 0.01, 0.05, d(g), dhh, d(f9), d(t8), d(Kf))
        P, Q = oracle.face_frame(*x[:3], *X[:3])
        _, fm, Km = oracle.compute_membrane(*x[:3], *X[:3], 50.0, 0.01, P, Q)
        _, fi, Mi = oracle.compute_inertial(*x[:3], *X[:3], g, 0.05)
        ref = Mi + dhh * Km
        for b, (i, j) in enumerate(FACE_PAIRS):
            worstF = max(worstF, np.abs(Kf[9 * b:9 * b + 9].reshape(3, 3) - ref[3 * i:3 * i + 3, 3 * j:3 * j + 3]).max() / np.abs(ref).max())
        # forces are ill-conditioned in the positions (strain = R F - I cancels): the error scale is stiffness x |x| x eps
        fs = max(np.abs(fm).max(), np.abs(fi).max()) + 1e-2 * np.abs(Km).max() * np.abs(x).max()
        worstf = max(worstf, np.abs(f9 - (fm + fi)).max() / fs)
        assert abs(t8[0] / 12 - Mi[0, 0]) <= 1e-15 * abs(Mi[0, 0])
    assert worstE < 1e-12 and worstF < 1e-12 and worstf < 1e-11, (worstE, worstF, worstf)
