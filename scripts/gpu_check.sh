#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms), ncu launch list + one full capture of the fill kernel.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "bench ref rc=$?"
timeout 300 python bench.py --workload ensemble64 --steps 10 --warmup 3 --no-cpu --no-cd > $OUT/bench_ens.json 2> $OUT/bench_ens.err; echo "ens rc=$?"
timeout 300 python bench.py --workload sheet256 --steps 50 --warmup 5 --no-cpu --no-cd > $OUT/bench_256.json 2> $OUT/bench_256.err; echo "256 rc=$?"
EOLC_FORCES_PIPELINE=rows timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-cd > $OUT/bench_rows.json 2> $OUT/bench_rows.err; echo "rows A/B rc=$?"
# launch list of the same bench command (cold-cache, serialised: shares only)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
# full capture of the fill kernel
timeout 400 ncu --set full --clock-control none --import-source on -k regex:assemble_ -s 3 -c 1 -o $OUT/prof_fill \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-cd > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -c 1500 $OUT/bench.json
