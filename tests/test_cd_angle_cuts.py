"""Section C of the narrow phase calls acos three times and only COMPARES the results with thresholds
(boxTriCollision.cpp:879-887, :907-915).  The CUDA path has no device acos: it compares the cosine with critical doubles the host
derives from ITS OWN libm (csrc/cd.cu first_true / eolc_cd_angle_cuts).  This test checks on the CPU that those cuts reproduce the
libm-acos decisions exactly — for every double within +-3000 ulps of each switch, and for random arguments (math.acos IS libm's)."""
import math

import numpy as np

import eol_cloth_b200 as E
from eol_cloth_b200 import capi

T = 2.0 * math.pi / 180.0


def _acos(c):
    return math.acos(c) if -1.0 <= c <= 1.0 else math.nan  # libm returns NaN outside the domain; every comparison is then false


def _ref_parallel(c):      # :880-881 / :886-887
    a = _acos(c)
    return abs(a) < T or abs(math.pi - a) < T


def _ref_wedge(c, angleCD):  # :906-915
    angleCN = _acos(c)
    if angleCD < 0.0:
        angleCD, angleCN = -angleCD, -angleCN
    return angleCN < -T or angleCN - angleCD > T


def _cut_parallel(c, hi, lo):
    return (hi <= c <= 1.0) or (-1.0 <= c <= lo)


def _cut_wedge(c, cw):
    return -1.0 <= c <= 1.0 and c < cw


def _around(c0, n=3000):
    out = [c0]
    up = dn = c0
    for _ in range(n):
        up = math.nextafter(up, 2.0)
        dn = math.nextafter(dn, -2.0)
        out += [up, dn]
    return out


def _cuts(rot=None, whd=(1.2, 1.5, 1.0)):
    cuts = np.zeros(14)
    Em = np.ascontiguousarray(E.meshgen.box_frame(np.array([0.3, 0.2, 0.1]), rot).reshape(16))
    capi.check(capi.lib().eolc_cd_angle_cuts(capi.dptr(np.array(whd, float)), capi.dptr(Em), capi.dptr(cuts)))
    return cuts


def test_parallel_cuts_reproduce_libm_decisions():
    cuts = _cuts()
    hi, lo = cuts[0], cuts[1]
    assert abs(hi - math.cos(T)) < 1e-15 and abs(lo + math.cos(T)) < 1e-15
    rng = np.random.default_rng(0)
    samples = _around(hi) + _around(lo) + [-1.0, 1.0, 0.0, -0.0, math.nextafter(1.0, 2.0), math.nextafter(-1.0, -2.0), 1.5, -7.0,
                                           math.nan, math.inf] + rng.uniform(-1.001, 1.001, 20000).tolist()
    for c in samples:
        assert _cut_parallel(c, hi, lo) == _ref_parallel(c), c


def test_wedge_cuts_reproduce_libm_decisions():
    ax = np.array([0.3, -0.5, 0.8]) / np.linalg.norm([0.3, -0.5, 0.8])
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(3) + math.sin(0.7) * K + (1 - math.cos(0.7)) * K @ K
    rng = np.random.default_rng(1)
    for rot in (None, R):
        cuts = _cuts(rot)
        # a box's adjacent faces are perpendicular: angleCD = acos(n1c.n1d) with n1c.n1d = 0 up to rounding of the rotated normals
        for k in range(12):
            cw = cuts[2 + k]
            assert abs(cw - math.cos(math.pi / 2 + T)) < 1e-12
        # the decision depends on angleCD, which the library computes per box edge; reproduce it from the cut itself:
        # the cut is the first c NOT rejected, so angleCD is recovered by checking both neighbours against every candidate
        # angleCD in a small set around pi/2 (here: exact check with the axis-aligned box, where n1c.n1d == 0 exactly)
    cuts = _cuts(None)
    angleCD = math.acos(0.0)
    for k in range(12):
        cw = cuts[2 + k]
        for c in _around(cw) + [-1.0, 1.0, 0.0, 1.0000001, -1.0000001, math.nan] + rng.uniform(-1.001, 1.001, 3000).tolist():
            assert _cut_wedge(c, cw) == _ref_wedge(c, angleCD), (k, c)
