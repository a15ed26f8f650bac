#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/d_relax.log
python scripts/time_relax2.py >> gpurun_out/d_relax.log 2>&1
for f in variants/libdbspm_b200_*.so; do
  DBSPM_B200_LIB=$PWD/$f timeout 300 python scripts/time_relax2.py >> gpurun_out/d_relax.log 2>&1
done
grep RELAX gpurun_out/d_relax.log
