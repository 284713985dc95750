#!/bin/bash
# Builds scratch/variants/libeolc_<name>.so: forces.cu recompiled with extra flags, linked with the default ctx.o / cd.o.
# Usage: bash scripts/build_variant.sh <name> "<extra nvcc flags>"     (run `make -C eol_cloth_b200/csrc` first)
set -e
cd "$(dirname "$0")/.."
NAME=$1; EXTRA=$2
C=eol_cloth_b200/csrc
mkdir -p scratch/variants scratch/obj
nvcc $EXTRA -O3 -std=c++17 -Xcompiler -ffp-contract=off -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v \
     -c ${SRC:-$C/forces.cu} -o scratch/obj/forces_$NAME.o 2> scratch/obj/forces_$NAME.ptxas.log || { cat scratch/obj/forces_$NAME.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o scratch/variants/libeolc_$NAME.so $C/ctx.o scratch/obj/forces_$NAME.o $C/cd.o -lcudart_static -ldl -lrt -lpthread
grep -A2 "assemble_tiles" scratch/obj/forces_$NAME.ptxas.log | grep -E "registers|spill" | head -3
