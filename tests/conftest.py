import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure). Built on demand; oracle/_ref is prebuilt where /root/reference is absent."""
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def hostmath():
    """g++ build of the SAME element arithmetic the CUDA kernels use (csrc/elements.cuh, csrc/cd_math.cuh)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "hostmath")
    so = os.path.join(d, "libhostmath.so")
    src = os.path.join(d, "hostmath.cpp")
    deps = [src, os.path.join(ROOT, "eol_cloth_b200", "csrc", "elements.cuh"), os.path.join(ROOT, "eol_cloth_b200", "csrc", "cd_math.cuh"), os.path.join(ROOT, "eol_cloth_b200", "csrc", "forces_plan.h"), os.path.join(ROOT, "eol_cloth_b200", "csrc", "tile_exec.cuh"), os.path.join(ROOT, "eol_cloth_b200", "csrc", "forces_eol.h")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=c++14", "-pthread", "-o", so, src])
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def ctx():
    import eol_cloth_b200 as E
    c = E.Context(0)
    yield c
    c.close()
