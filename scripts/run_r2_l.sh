#!/bin/bash
# round 2, call L: pipelined axis pass A/B + relax warp-footprint sweep
mkdir -p gpurun_out
timeout 600 python scripts/time_corr3.py large_slab > gpurun_out/l_corr.log 2>&1
grep CORR3 gpurun_out/l_corr.log
: > gpurun_out/l_relax.log
for fx in 1 2 4 8 16 32; do
  DBSPM_RELAX_FOOT_X=$fx timeout 300 python scripts/time_relax2.py >> gpurun_out/l_relax.log 2>&1
  echo "  ^ foot_x=$fx" >> gpurun_out/l_relax.log
done
grep "RELAX\|foot" gpurun_out/l_relax.log
