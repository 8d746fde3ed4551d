"""Developer tool: executed-instruction counts per SASS basic block of the scoring kernel from an .ncu-rep
(ncu --set full --import-source on). usage: ncu_sass.py rep n_ligands [min_instr_per_ligand]"""
import csv
import subprocess
import sys

rep, nlig = sys.argv[1], float(sys.argv[2])
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 500.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
iS, iI, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
data = [(r[iS].strip(), int(r[iI] or 0), int(r[iT] or 0)) for r in rows[2:] if len(r) > iI]
tot, tots = sum(d[1] for d in data), sum(d[2] for d in data)
print(f"instructions per ligand {tot / nlig:.0f}, stall samples {tots}")
i = 0
while i < len(data):
    j = i
    while j < len(data) and data[j][1] == data[i][1]:
        j += 1
    ins = (j - i) * data[i][1] / nlig
    if ins >= thr:
        smp = sum(d[2] for d in data[i:j])
        ops = " ".join(sorted({d[0].split()[0 if not d[0].startswith("@") else 1].split(".")[0] for d in data[i:j]}))
        print(f"[{i:5d}-{j - 1:5d}] n={j - i:3d} exec/lig={data[i][1] / nlig:8.1f} instr/lig={ins:8.0f} "
              f"({100 * ins * nlig / tot:4.1f}%) samples={100 * smp / tots:4.1f}%  {ops[:110]}")
    i = j
