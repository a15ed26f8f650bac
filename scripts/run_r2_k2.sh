#!/bin/bash
# round 2, call K2: single-GPU evidence -- bench line, reference arm, ncu launch list, ncu --set full of one step.
# The .ncu-rep files stay on the box (/tmp): only text summaries travel back (gpurun_out is capped at 64 MiB).
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || exit 1
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_reference_arm.err; echo "reference rc=$?"
K='regex:axis_pass|axis_tma|zfused|relax_|accumulate|compact|transpose|quad_yz'
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches_large_slab.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 48 -c 16 -o /tmp/r2_full -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-compact --no-cufft > gpurun_out/k_ncu_f.log 2>&1
python scripts/ncu_summary.py /tmp/r2_full.ncu-rep > gpurun_out/r2_ncu_full_large_slab_summary.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_full.ncu-rep relax_kernel 60 > gpurun_out/r2_ncu_lines_relax.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_full.ncu-rep zfused2 40 > gpurun_out/r2_ncu_lines_zfused.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_full.ncu-rep axis_pass2 40 > gpurun_out/r2_ncu_lines_axis.txt 2>&1
ncu -i /tmp/r2_full.ncu-rep --page raw --csv > /tmp/r2_raw.csv 2>/dev/null
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/r2_raw.csv')))
hdr = rows[0]
keep = [i for i, h in enumerate(hdr) if h in ('ID', 'Kernel Name') or any(k in h for k in ('gpu__time_duration', 'dram__bytes', 'warp_issue_stalled', 'pct_of_peak_sustained', 'hit_rate', 'inst_executed.sum', 'registers_per_thread', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld', 'l1tex__data_bank_conflicts', 'l1tex__data_pipe_lsu_wavefronts', 'thread_inst_executed_per_inst', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes'))]
with open('gpurun_out/r2_ncu_full_large_slab_metrics.csv', 'w', newline='') as fh:
    w = csv.writer(fh)
    for r in rows:
        w.writerow([r[i] for i in keep])
PY
python scripts/prof_direct.py > gpurun_out/k_direct_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:direct_kernel -s 1 -c 1 -o /tmp/r2_direct -f python scripts/prof_direct.py > gpurun_out/k_ncu_d.log 2>&1
python scripts/ncu_summary.py /tmp/r2_direct.ncu-rep > gpurun_out/r2_ncu_direct_summary.txt 2>&1
python scripts/ncu_lines.py /tmp/r2_direct.ncu-rep direct_kernel 25 > gpurun_out/r2_ncu_lines_direct.txt 2>&1
du -sh gpurun_out; tail -2 gpurun_out/r2_bench_n1.err
