#!/bin/bash
# round 2, call K: single-GPU evidence -- bench line, ncu launch list, ncu --set full of one step (each after the plain run exited 0)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || exit 1
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference rc=$?"
K='regex:axis_pass|axis_tma|zfused|relax_|accumulate|compact|transpose|quad_yz'
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches_large_slab.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 48 -c 16 -o gpurun_out/r2_full_large_slab -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_ncu_f.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_full_large_slab.ncu-rep > gpurun_out/r2_ncu_full_large_slab_summary.txt 2>&1
python scripts/prof_direct.py > gpurun_out/k_direct_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:direct_kernel -s 1 -c 1 -o gpurun_out/r2_direct -f python scripts/prof_direct.py > gpurun_out/k_ncu_d.log 2>&1
python scripts/ncu_summary.py gpurun_out/r2_direct.ncu-rep > gpurun_out/r2_ncu_direct_summary.txt 2>&1
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out
cat gpurun_out/r2_ncu_full_large_slab_summary.txt gpurun_out/r2_ncu_direct_summary.txt | cut -c1-400
tail -2 gpurun_out/r2_bench_n1.err; cut -c1-700 gpurun_out/r2_bench_reference_arm.json
