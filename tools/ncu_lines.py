"""Developer tool: per-source-line instruction / stall-sample shares from an .ncu-rep (needs -lineinfo)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "dram__bytes_read.sum",
          "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"):
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:60s} {vals[i]} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[2]
ix_inst = hdr.index("Instructions Executed")
ix_s = hdr.index("Warp Stall Sampling (All Samples)")
lines = []
for r in rows[3:]:
    if len(r) < len(hdr) or r[0] == "":
        continue
    try:
        lines.append((int(r[0]), r[1], int(r[ix_inst] or 0), int(r[ix_s] or 0)))
    except ValueError:
        pass
tot = sum(l[2] for l in lines)
tots = sum(l[3] for l in lines)
print("total inst", tot, "samples", tots)
for l in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{l[0]:5d} inst {100*l[2]/tot:5.1f}% samp {100*l[3]/tots:5.1f}%  {l[1].strip()[:105]}")
