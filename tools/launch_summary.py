"""Developer tool: per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: launch_summary.py launches.csv "command" > summary.txt"""
import csv
import sys
from collections import defaultdict

path, cmd = sys.argv[1], sys.argv[2]
rows = []
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.reader(lines):
    rows.append(r)
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = defaultdict(float), defaultdict(int)
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    except (ValueError, IndexError):
        continue
    tot[r[ik]] += v
    cnt[r[ik]] += 1
s = sum(tot.values())
print(cmd)
print(f"{len(rows) - 1} launches, {s:.3f} ms of GPU time in total (cold-cache, serialised by ncu: compare shares, not absolutes)")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{v:12.3f} ms total {cnt[k]:5d} launches {100 * v / s:6.2f}%  {k[:110]}")
