#!/bin/bash
# Iteration loop for the Forces::fill kernel: parity tests, memcheck of a small fill, bench of the 1024^2 sheet, one ncu capture.
# Usage: bash scripts/gpu_quick.sh <tag> [ncu:0|1] [sanitizer:0|1]
TAG=${1:-q}; NCU=${2:-1}; SAN=${3:-0}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python -m pytest tests/test_forces_gpu.py -x -q > $OUT/pytest_forces.log 2>&1; echo "pytest forces rc=$?"; tail -3 $OUT/pytest_forces.log
if [ "$SAN" = "1" ]; then
  timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_forces_gpu.py -x -q -k "golden or shuffled or empty or phases" > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $OUT/memcheck.log
  timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_forces_gpu.py -x -q -k "golden" > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 $OUT/racecheck.log
fi
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu --no-cd > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench.json")); print("ms/fill", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e: print("bench parse failed", e); print(open("$OUT/bench.err").read()[-2000:])
PY
timeout 240 python bench.py --workload ensemble64 --steps 10 --warmup 3 --no-cpu --no-cd > $OUT/bench_ens.json 2> $OUT/bench_ens.err; echo "ens rc=$?"
timeout 240 python bench.py --workload sheet256 --steps 50 --warmup 5 --no-cpu --no-cd > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "256 rc=$?"
python - <<PY
import json
for n in ("bench_ens","bench_256"):
    try:
        d=json.load(open("$OUT/%s.json"%n)); print(n, "ms", d["ms_per_step"], "value", d["value"], "frac", d["roofline"]["frac"])
    except Exception as e: print(n, "failed", e)
PY
if [ "$NCU" = "1" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:assemble_ -s 3 -c 1 -o $OUT/prof_fill \
      python bench.py --steps 3 --warmup 3 --no-cpu --no-cd > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
