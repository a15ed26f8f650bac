"""Developer aid: stall reasons (cycles per issue), L1 / L2 hit rates and wavefront counts per kernel of an ncu report.
usage: ncu_stalls.py report.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
kn = hdr.index("Kernel Name")
sel = sys.argv[2] if len(sys.argv) > 2 else ""
st = [i for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
extra = ["gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
         "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
         "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
         "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
         "dram__bytes_read.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum"]
for r in rows[2:]:
    if sel not in r[kn]:
        continue
    print("==", r[kn][:70])
    vals = []
    for i in st:
        try:
            vals.append((float(r[i]), hdr[i].replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
    print("   stalls/issue:", ", ".join(f"{h}={v:.2f}" for v, h in sorted(vals, reverse=True)[:9]))
    for k in extra:
        if k in hdr:
            print(f"   {k} = {r[hdr.index(k)]}")
