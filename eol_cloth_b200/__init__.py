"""eol_cloth_b200 — B200-native (sm_100a) hot path of sueda/eol-cloth behind a C ABI.

Only the per-step force/Jacobian assembly (`Forces::fill`) and the collision narrow phase (`CD`/`CD2`) live here;
see DESIGN.md.  The compute lives in `libeolc_b200.so` (CUDA, built by `eol_cloth_b200/csrc/Makefile`); this
package is the thin host-side mirror of the reference's interface used by the tests and the benchmark.
There is NO CPU fallback: importing works without a GPU, every compute call raises `EolcError` without one.
"""
from .capi import EolcError, Context, device_count, lib_path  # noqa: F401
from .forces import Forces, ForcesPlan, Material  # noqa: F401
from .collisions import CD, CD2, CollisionPlan, Obstacles, CONTACT_DTYPE, contact_rows  # noqa: F401
from . import meshgen  # noqa: F401
