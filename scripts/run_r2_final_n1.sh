#!/bin/bash
# round 2, single-GPU evidence of the end-of-round build: bench line, reference arm, sweep16 line, ncu launch list,
# ncu --set full of one step (each ncu pass after the same command exited 0 without ncu).  The .ncu-rep files stay on the
# box (/tmp): only text summaries travel back (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || exit 1
python bench.py --workload sweep16 --steps 3 --warmup 3 > gpurun_out/r2_bench_sweep16_n1.json 2> gpurun_out/r2_bench_sweep16_n1.err; echo "sweep16 rc=$?"
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference rc=$?"
K='regex:axis_pass|axis_tma|zfused|relax_|accumulate|compact|transpose|quad_yz'
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches_large_slab.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 48 -c 16 -o /tmp/r2_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_ncu_f.log 2>&1
python scripts/ncu_summary.py /tmp/r2_full.ncu-rep > gpurun_out/r2_ncu_full_large_slab_summary.txt 2>&1
python scripts/ncu_stalls.py /tmp/r2_full.ncu-rep > gpurun_out/r2_ncu_full_large_slab_stalls.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_full.ncu-rep relax_kernel 60 > gpurun_out/r2_ncu_lines_relax.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_full.ncu-rep zfused2 40 > gpurun_out/r2_ncu_lines_zfused.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_full.ncu-rep axis_pass2 40 > gpurun_out/r2_ncu_lines_axis.txt 2>&1
du -sh gpurun_out; tail -2 gpurun_out/r2_bench_n1.err; cut -c1-300 gpurun_out/r2_bench_sweep16_n1.json
