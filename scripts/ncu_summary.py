"""Developer aid: one-line-per-kernel summary of an ncu report (the metrics DESIGN/profiles quote).
usage: ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "ms"), ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem%"),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankconf"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wf"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__occupancy_limit_shared_mem", "lim_smem"), ("launch__occupancy_limit_registers", "lim_regs"),
        ("smsp__inst_executed.sum", "inst"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%")]
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0][-40:]
    parts = []
    for k, lab in want:
        if k in ix:
            v = r[ix[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            parts.append(f"{lab}={v}{units[ix[k]] if lab in ('rd','wr') else ''}")
    print(name, " ".join(parts))
