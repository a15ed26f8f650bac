"""Developer aid: aggregate an ncu report's source page by CUDA source line.
usage: ncu_lines.py report.ncu-rep kernel_regex [topN]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "-k", f"regex:{rx}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; agg = []; cur = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = {h: i for i, h in reversed(list(enumerate(r)))}; continue
    if hdr is None or len(r) < 8 or r[0] in ("Function Name",): continue
    if r[0] != "" and r[2] == "-":
        try:
            agg.append((cur, int(r[0]), r[1].strip()[:80], int(r[hdr["# Samples"]]), int(r[hdr["Instructions Executed"]]),
                        int(r[hdr["Thread Instructions Executed"]])))
        except ValueError:
            pass
ts = sum(a[3] for a in agg) or 1; ti = sum(a[4] for a in agg) or 1
print(f"total samples {ts} warp-inst {ti} thread-inst/warp-inst {sum(a[5] for a in agg)/ti:.1f}")
for f, ln, src, s, i, t in sorted(agg, key=lambda a: -a[4])[:top]:
    print(f"{f}:{ln:4d} inst {100*i/ti:5.1f}% ({i/1e6:8.1f}M) samp {100*s/ts:5.1f}% thr/w {t/max(i,1):5.1f} | {src}")
