#!/bin/bash
# round 2, call B: ncu --set full of the relax kernel, quad layout and y-fastest layout
mkdir -p gpurun_out
python scripts/time_relax2.py > gpurun_out/b_time.log 2>&1
RELAX_LAYOUT=2 python scripts/prof_relax.py > gpurun_out/b_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:relax_kernel -s 1 -c 1 -f -o gpurun_out/r2_relax_quad python scripts/prof_relax.py > gpurun_out/b_ncu2.log 2>&1
RELAX_LAYOUT=1 python scripts/prof_relax.py > gpurun_out/b_plain1.log 2>&1 &&
RELAX_LAYOUT=1 ncu --set full --clock-control none --import-source on -k regex:relax_kernel -s 1 -c 1 -f -o gpurun_out/r2_relax_l1 python scripts/prof_relax.py > gpurun_out/b_ncu1.log 2>&1
tail -3 gpurun_out/b_time.log gpurun_out/b_ncu2.log gpurun_out/b_ncu1.log
ls -la gpurun_out/*.ncu-rep
