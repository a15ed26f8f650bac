#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/e_relax.log
for hpl in 24 1 2 4 8; do
  echo "== HPL $hpl" >> gpurun_out/e_relax.log
  DBSPM_RELAX_HEIGHTS_PER_LAUNCH=$hpl python scripts/time_relax2.py >> gpurun_out/e_relax.log 2>&1
  DBSPM_RELAX_HEIGHTS_PER_LAUNCH=$hpl DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_t32mb16.so python scripts/time_relax2.py >> gpurun_out/e_relax.log 2>&1
done
for hpl in 1 2; do
  echo "== HPL $hpl layout 1 / t64" >> gpurun_out/e_relax.log
  DBSPM_RELAX_HEIGHTS_PER_LAUNCH=$hpl RELAX_LAYOUT=1 python scripts/time_relax2.py >> gpurun_out/e_relax.log 2>&1
  DBSPM_RELAX_HEIGHTS_PER_LAUNCH=$hpl DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_t64mb8.so python scripts/time_relax2.py >> gpurun_out/e_relax.log 2>&1
done
grep "RELAX\|==" gpurun_out/e_relax.log
