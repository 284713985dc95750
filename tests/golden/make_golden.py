"""Generates tests/golden/*.npz — all of them from THE REFERENCE'S OWN CODE, compiled unmodified against oracle/mini_eigen by
oracle/Makefile; `source` in each file names the library:
  forces_* / normals_*: oracle/_ref/libforces_ref.so = Forces.cpp + UtilEOL.cpp + conversions.cpp + Compute*.cpp + ArcSim's mesh /
                        geometry / util / vectors / transformation .cpp  (Forces::fill on a mesh built like Cloth::build; compute_ws_data);
  cd_*:                 oracle/_ref/libbtc_ref.so = boxTriCollision.cpp + Collisions.cpp + raytri.cpp.

Run in the build container (where /root/reference exists):  python tests/golden/make_golden.py
The fixtures pin (a) the oracle against silent drift and (b) the CUDA path on the GPU box, where the reference
sources are absent.  Inputs are regenerated from seeds by the tests; only outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
import eol_cloth_b200 as E  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def eol_line(n):
    """EoL nodes of the EOL fixtures: the interior nodes of the grid line j = n // 2, EoL_index in grid order."""
    eol = np.full(n * n, -1, np.int32)
    line = np.arange(1, n - 1) * n + n // 2
    eol[line] = np.arange(line.size)
    return eol


def forces_case(gen, n, seed, eol=None):
    X, fn = getattr(E.meshgen, gen)(n)
    es = E.meshgen.edge_stencils(X.shape[0], fn)
    x = E.meshgen.drape_state(X, seed=seed)
    r = O.ref_forces_fill(fn, x, X, eol_index=eol)
    assert np.array_equal(r["edge_stencil"], es)
    return dict(source=np.array("libforces_ref"), f=r["f"], M_outer=r["M"][0], M_inner=r["M"][1], M_vals=r["M"][2], K_outer=r["MDK"][0],
                K_inner=r["MDK"][1], K_vals=r["MDK"][2])


def cd_case(gen, n, centre, seed, rot=None, points=False):
    X, fn = getattr(E.meshgen, gen)(n)
    x = E.meshgen.box_scene_state(X, seed=seed, centre=np.asarray(centre))
    pxyz = pn = None
    if points:
        pxyz = np.array([[0.25, 0.25, x[:, 2].max() - 4e-3], [0.1, 0.8, -0.2], x[5] + 1e-3])
        pn = np.array([[0, 0, 1.0], [0, 0, 1.0], [0, 0, 1.0]])
    out = {"source": np.array("libbtc_ref")}
    for name, which in (("cd", 1), ("cd2", 0)):
        out[name] = O.ref_cd(fn, x, E.meshgen.BOX_THRESHOLD, pxyz, pn, E.meshgen.BOX_WHD[None], E.meshgen.box_frame(centre, rot)[None], which)
    return out


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "forces_regular2_n12.npz"), **forces_case("regular2", 12, 0))
    np.savez_compressed(os.path.join(HERE, "forces_build4_n7.npz"), **forces_case("build4", 7, 1))
    np.savez_compressed(os.path.join(HERE, "forces_eol_regular2_n12.npz"), **forces_case("regular2", 12, 2, eol_line(12)))
    _, fn_, nn_ = O.ref_mesh_data(E.meshgen.build4(7)[1], E.meshgen.drape_state(E.meshgen.build4(7)[0], seed=1), E.meshgen.build4(7)[0])
    np.savez_compressed(os.path.join(HERE, "normals_build4_n7.npz"), source=np.array("libforces_ref"), face_n=fn_, node_n=nn_)
    np.savez_compressed(os.path.join(HERE, "cd_regular2_n24.npz"), **cd_case("regular2", 24, E.meshgen.BOX_CENTRE, 0))
    c3b = np.array([0.9175, -0.25, -0.549])
    np.savez_compressed(os.path.join(HERE, "cd_build4_n16_corner.npz"), **cd_case("build4", 16, c3b, 1, points=True))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
