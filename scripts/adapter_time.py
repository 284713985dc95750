"""Developer tool: step time of Forces::fill through the EXECUTED adapter (ArcSim pointer mesh in, Eigen members out), copy and zero-copy modes.
usage: python scripts/adapter_time.py n"""
import sys; sys.path.insert(0,'/root/repo')
from oracle import oracle as O
import eol_cloth_b200 as E
n = int(sys.argv[1])
X, fn = E.meshgen.regular2(n)
for mode in (True, "zero_copy"):
    first, steady = O.ref_forces_step_seconds(fn, E.meshgen.drape_state(X, seed=0), X, steps=5, adapter=mode)
    print(f"adapter ({mode}) n={n}: first {first*1e3:.1f} ms, steady {steady*1e3:.2f} ms")
