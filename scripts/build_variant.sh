#!/bin/bash
# Developer aid: build a variant of the CUDA library with extra -D flags for relax.cu / corr3d.cu
#   scripts/build_variant.sh NAME "<relax flags>" "<corr3d flags>"  -> gpurun_variants/libdbspm_b200_NAME.so
set -e
cd "$(dirname "$0")/../dbspm_b200/csrc"
OUT=../../variants; mkdir -p $OUT
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC $ARCH"
T=$(mktemp -d)
nvcc $COMMON $3 -Xptxas -v -c corr3d.cu -o $T/corr3d.o 2> $OUT/$1.corr3d.ptxas.log
nvcc $COMMON --fmad=false $2 -Xptxas -v -c relax.cu -o $T/relax.o 2> $OUT/$1.relax.ptxas.log
nvcc $COMMON -c capi.cu -o $T/capi.o
nvcc $COMMON -x cu -c plan.cpp -o $T/plan.o
nvcc $ARCH -shared -o $OUT/libdbspm_b200_$1.so $T/corr3d.o $T/relax.o $T/capi.o $T/plan.o
rm -rf $T
grep -A2 "relax_kernel" $OUT/$1.relax.ptxas.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $1"
