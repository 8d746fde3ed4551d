"""Developer tool: one JSON record per kernel launch of an `ncu --set full` report (CNN kernels, profiles/*.json).
usage: ncu_kernels.py rep out.json "command" """
import csv
import json
import subprocess
import sys

rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = (
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
)
tensor_like = [h for h in hdr if "tensor" in h and "pct_of_peak_sustained_active" in h and h.startswith("sm__")]
recs = []
for v in rows[2:]:
    if len(v) < len(hdr):
        continue
    r = {"kernel": v[hdr.index("Kernel Name")], "id": v[hdr.index("ID")]}
    for k in list(keep) + tensor_like:
        if k in hdr:
            r[k] = [v[hdr.index(k)], units[hdr.index(k)]]
    recs.append(r)
with open(out, "w") as f:
    json.dump({"command": cmd, "launches": recs}, f, indent=1)
for r in recs:
    t = r.get("gpu__time_duration.sum", ["?", ""])
    tp = [r[k][0] for k in tensor_like if k in r][:2]
    print(f"{r['kernel'][:60]:60s} {t[0]:>10s} {t[1]:5s} dram R/W {r.get('dram__bytes_read.sum', ['?'])[0]}/{r.get('dram__bytes_write.sum', ['?'])[0]} tensor% {tp}")
