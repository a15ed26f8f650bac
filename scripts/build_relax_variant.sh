#!/bin/bash
# Developer aid: variant of the library that differs in relax.cu only (reuses the objects of the default build)
#   scripts/build_relax_variant.sh NAME "<relax flags>"  -> variants/libdbspm_b200_NAME.so
set -e
cd "$(dirname "$0")/../dbspm_b200/csrc"
OUT=../../variants; mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
T=$(mktemp -d)
nvcc $COMMON --fmad=false $2 -Xptxas -v -c relax.cu -o $T/relax.o 2> $OUT/$1.relax.ptxas.log
nvcc $ARCH -shared -o $OUT/libdbspm_b200_$1.so corr3d.o $T/relax.o capi.o plan.o
rm -rf $T
grep -A2 "relax_kernelILi2" $OUT/$1.relax.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1"
