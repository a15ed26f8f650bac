#!/bin/bash
# round 2, call W: occupancy variants of the relax kernel again, now with the 8 x 4 patch and the dynamic schedule
mkdir -p gpurun_out
: > gpurun_out/w_relax.log
for v in "" s0mb20 s1mb20 s2mb20 s2mb24; do
  if [ -n "$v" ]; then export DBSPM_B200_LIB=$PWD/variants/libdbspm_b200_$v.so; else unset DBSPM_B200_LIB; fi
  timeout 300 python scripts/time_relax2.py >> gpurun_out/w_relax.log 2>&1
done
grep RELAX gpurun_out/w_relax.log
