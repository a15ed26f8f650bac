#!/bin/bash
# round 2, call M: radix-8 axis plans with three resident CTAs; relax drift with the 8 x 4 warp footprint
mkdir -p gpurun_out
: > gpurun_out/m_corr.log
for v in "" r8mb3 r8mb2; do
  if [ -n "$v" ]; then export DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_$v.so; else unset DBSPM_B200_LIB; fi
  echo "== lib ${v:-default}" >> gpurun_out/m_corr.log
  timeout 600 python scripts/time_corr3.py large_slab >> gpurun_out/m_corr.log 2>&1
done
grep "CORR3\|==" gpurun_out/m_corr.log | grep -v pipe
: > gpurun_out/m_relax.log
for v in "" drift1 drift2 drift24; do
  if [ -n "$v" ]; then export DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_$v.so; else unset DBSPM_B200_LIB; fi
  timeout 300 python scripts/time_relax2.py >> gpurun_out/m_relax.log 2>&1
done
unset DBSPM_B200_LIB
for fx in 4 8; do for hpl in 1 4; do
  echo "== foot_x=$fx hpl=$hpl" >> gpurun_out/m_relax.log
  DBSPM_RELAX_FOOT_X=$fx DBSPM_RELAX_HEIGHTS_PER_LAUNCH=$hpl timeout 300 python scripts/time_relax2.py >> gpurun_out/m_relax.log 2>&1
done; done
grep "RELAX\|==" gpurun_out/m_relax.log
