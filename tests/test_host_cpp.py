"""The C++ host layer above the C ABI (include/eolc_host.hpp), driven from a C++ program with an ArcSim-shaped pointer mesh —
the code path the reference-side adapters (adapter/*.cpp) take.  CPU: it builds, links against libeolc_b200.so and, without a
GPU, fails the way the reference fails (message + abort()).  GPU: results equal the oracle's."""
import os
import struct
import subprocess

import numpy as np
import pytest

import eol_cloth_b200 as E
from util import assert_close_tol, assert_contacts_equal, block_row_scale

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_driver.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_driver")
MAT = E.Material.DEFAULT
GRAV = (0.0, 0.0, -9.8)
H = 0.5e-2


@pytest.fixture(scope="module")
def driver():
    libdir = os.path.join(ROOT, "eol_cloth_b200")
    deps = [SRC, os.path.join(ROOT, "include", "eolc_host.hpp"), os.path.join(ROOT, "include", "eolc.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(p) > os.path.getmtime(EXE) for p in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++11", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"), SRC, "-o", EXE,
                               "-L", libdir, "-leolc_b200", "-Wl,-rpath," + libdir])
    return EXE


def _write_input(path, X, fn, x, centre):
    par = list(MAT) + list(GRAV) + [H, E.meshgen.BOX_THRESHOLD] + list(E.meshgen.BOX_WHD) + list(E.meshgen.box_frame(centre).reshape(-1))
    with open(path, "wb") as f:
        f.write(struct.pack("ii", X.shape[0], fn.shape[0]))
        f.write(np.ascontiguousarray(x, np.float64).tobytes()); f.write(np.ascontiguousarray(X, np.float64).tobytes())
        f.write(np.ascontiguousarray(fn, np.int32).tobytes()); f.write(np.array(par, np.float64).tobytes())


def test_host_cpp_aborts_without_gpu(driver, tmp_path):
    if E.device_count() > 0:
        pytest.skip("a CUDA device is present")
    X, fn = E.meshgen.regular2(4)
    _write_input(tmp_path / "in.bin", X, fn, E.meshgen.drape_state(X, seed=0), np.array([0.9175, -0.25, -0.549]))
    r = subprocess.run([driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert r.returncode != 0                       # abort(), like the reference's error paths
    assert "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout
    assert not os.path.exists(tmp_path / "out.bin")


@pytest.mark.gpu
@pytest.mark.parametrize("with_eol", [False, True], ids=["lagrangian", "eol"])
def test_host_cpp_matches_oracle(driver, oracle, tmp_path, with_eol):
    X, fn = E.meshgen.regular2(24)
    c = np.array([0.9175, -0.25, -0.549])
    x = E.meshgen.box_scene_state(X, seed=3, centre=c)
    _write_input(tmp_path / "in.bin", X, fn, x, c)
    eol = None
    if with_eol:      # the driver flags the same nodes on its pointer mesh (Node::EoL / EoL_index), flatten() carries them over
        eol = np.full(X.shape[0], -1, np.int32)
        line = np.arange(1, 23) * 24 + 12
        eol[line] = np.arange(line.size)
    r = subprocess.run([driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")] + (["24"] if with_eol else []), capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    raw = open(tmp_path / "out.bin", "rb").read()
    dof, nE, cutoff, ncls, nnzM, nnzK, m_updated_second, _ = struct.unpack_from("iiiiqqii", raw, 0)
    off = struct.calcsize("iiiiqqii")
    # second fill on unchanged X: M skipped on the Lagrangian mesh (EOLC_FILL_M_UNCHANGED), recomputed with EoL nodes
    assert m_updated_second == (1 if with_eol else 0)

    def take(dtype, n):
        nonlocal off
        a = np.frombuffer(raw, dtype=dtype, count=n, offset=off); off += a.nbytes
        return a
    es = take(np.int32, 4 * nE).reshape(-1, 4)
    f = take(np.float64, dof)
    Mo, Mi, Mv = take(np.int32, dof + 1), take(np.int32, nnzM), take(np.float64, nnzM)
    Ko, Ki, Kv = take(np.int32, dof + 1), take(np.int32, nnzK), take(np.float64, nnzK)
    cls = take(E.CONTACT_DTYPE, ncls)
    N = X.shape[0]
    assert cutoff == 3 * N and dof == 3 * N + (2 * 22 if with_eol else 0)
    assert np.array_equal(es, E.meshgen.edge_stencils(N, fn))          # flatten() == the generator's ArcSim edge order
    ref = oracle.forces_fill(fn, es, x, X, tuple(MAT), GRAV, H, eol_index=eol)
    assert_close_tol(f, ref["f"], np.abs(ref["f"]).max(), 1e-10, "f")
    for name, got in (("M", (Mo, Mi, Mv)), ("MDK", (Ko, Ki, Kv))):
        o, i, v = ref[name]
        assert np.array_equal(got[0], o) and np.array_equal(got[1], i), name
        assert_close_tol(got[2], v, block_row_scale(o, v, N), 1e-10, name)
    obs_E = E.meshgen.box_frame(c)[None]
    refc = oracle.cd(fn, x, E.meshgen.BOX_THRESHOLD, None, None, E.meshgen.BOX_WHD[None], obs_E, 0, 0)
    assert len(refc) > 0
    assert_contacts_equal(cls, refc, what="C++ host CD2")
