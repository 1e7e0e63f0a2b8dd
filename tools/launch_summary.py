"""Condenses an ncu launch list (gpu__time_duration.sum per launch) into per-kernel totals and shares.
usage: python tools/launch_summary.py launches.csv [out.md] [first_id [last_id]]   (only launches with first_id <= ID < last_id)"""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, gi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Metric Unit")
agg = defaultdict(list)
first_id = int(sys.argv[3]) if len(sys.argv) > 3 else 0
last_id = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 60
for r in data:
    if len(r) <= vi or not (first_id <= int(r[0]) < last_id):
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)
    name = r[ki].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
    agg[name].append(v)
tot = sum(sum(v) for v in agg.values())
lines = ["| kernel | launches | avg us | total ms | share |", "|---|---:|---:|---:|---:|"]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if sum(v) / tot < 0.002:
        continue
    lines.append("| `%s` | %d | %.1f | %.2f | %.1f %% |" % (k[:60], len(v), sum(v) / len(v), sum(v) / 1e3, 100 * sum(v) / tot))
lines.append("| total (%d launches) | | | %.2f | |" % (sum(len(v) for v in agg.values()), tot / 1e3))
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2 and sys.argv[2] != "-":
    open(sys.argv[2], "w").write(out + "\n")
