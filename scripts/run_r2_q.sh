#!/bin/bash
# round 2, call Q: ncu --set full of the relax kernel with the 8 x 4 warp patch (after the plain run exited 0)
mkdir -p gpurun_out
python scripts/time_relax2.py > gpurun_out/q_relax_plain.log 2>&1 || exit 1
grep RELAX gpurun_out/q_relax_plain.log
python scripts/prof_relax.py > gpurun_out/q_prof_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:relax_kernel -s 1 -c 1 -o /tmp/q_relax -f python scripts/prof_relax.py > gpurun_out/q_ncu.log 2>&1
python scripts/ncu_summary.py /tmp/q_relax.ncu-rep > gpurun_out/r2_ncu_relax_patch_summary.txt 2>&1
python scripts/ncu_stalls.py /tmp/q_relax.ncu-rep relax > gpurun_out/r2_ncu_relax_patch_stalls.txt 2>&1
python scripts/ncu_lines.py /tmp/q_relax.ncu-rep relax_kernel 50 > gpurun_out/r2_ncu_lines_relax_patch.txt 2>&1
cat gpurun_out/r2_ncu_relax_patch_summary.txt gpurun_out/r2_ncu_relax_patch_stalls.txt | cut -c1-600
