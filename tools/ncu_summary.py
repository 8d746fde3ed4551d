"""Developer tool: summarise an `ncu --set full` report of the scoring kernel into profiles/*.json.
usage: ncu_summary.py rep n_ligands out.json "command line used" [--latest]"""
import csv
import json
import subprocess
import sys

rep, nlig, out, cmd = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = (
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
)
m = {k: [vals[hdr.index(k)], units[hdr.index(k)]] for k in keep if k in hdr}
unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
dram = sum(float(m[k][0]) * unit[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
inst = float(m["smsp__inst_executed.sum"][0])
doc = {
    "command": cmd, "kernel": hdr and vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "pmnet_score_kernel",
    "ligands_in_launch": nlig, "dram_bytes_per_ligand": dram / nlig, "warp_instructions_per_ligand": inst / nlig,
    "metrics": m,
}
with open(out, "w") as f:
    json.dump(doc, f, indent=1)
if "--latest" in sys.argv:
    import os

    with open(os.path.join(os.path.dirname(out), "scoring_kernel_latest.json"), "w") as f:
        json.dump({"source": out[out.index("profiles"):] if "profiles" in out else out,
                   "dram_bytes_per_ligand": dram / nlig, "warp_instructions_per_ligand": inst / nlig,
                   "captured_ligands": nlig}, f, indent=1)
print(json.dumps({k: doc[k] for k in ("dram_bytes_per_ligand", "warp_instructions_per_ligand")}))
