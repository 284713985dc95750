#!/bin/bash
# Builds an alternative libeolc (A/B experiments): scripts/build_variant.sh <name> "<extra nvcc flags>"
set -e
cd "$(dirname "$0")/.."
mkdir -p scratch/variants
make -s -C eol_cloth_b200/csrc clean >/dev/null
make -s -C eol_cloth_b200/csrc EXTRA="$2" OUT="$(pwd)/scratch/variants/libeolc_$1.so"
grep -A3 "assemble_tiles" eol_cloth_b200/csrc/forces.ptxas.log | grep -E "Used|spill"
make -s -C eol_cloth_b200/csrc clean >/dev/null
