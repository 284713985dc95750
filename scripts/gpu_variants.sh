#!/bin/bash
# A/B alternative builds (scratch/variants/libeolc_<name>.so): parity test of the 256^2 sheet + bench of the 1024^2 sheet
OUT=gpurun_out/${1:-var}; mkdir -p $OUT
shift
for so in "$@"; do
  name=$(basename $so .so)
  if [ "$so" = "default" ]; then unset EOLC_LIB; name=default; else export EOLC_LIB=$(pwd)/$so; fi
  timeout 120 python -m pytest tests/test_forces_gpu.py -x -q -k "256 or golden or shuffled" > $OUT/pytest_$name.log 2>&1; echo "$name pytest rc=$? $(tail -1 $OUT/pytest_$name.log)"
  timeout 180 python bench.py --steps 20 --warmup 5 --no-cpu --no-cd > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json")); print("$name", "ms/fill %.4f"%d["ms_per_step"], "frac %.3f"%d["roofline"]["frac"])
except Exception as e: print("$name bench failed", e, open("$OUT/bench_$name.err").read()[-1500:])
PY
done
