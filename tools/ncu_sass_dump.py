"""Dump SASS of one kernel with per-instruction exec counts and stall samples: python tools/ncu_sass_dump.py rep kernel_index [min_exec]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], int(sys.argv[2]); mn = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
k = -1; hdr = None
for r in rows:
    if r and r[0] == "Kernel Name": k += 1; continue
    if r and r[0] == "Address": hdr = r; continue
    if k != kid or hdr is None or len(r) < len(hdr): continue
    ie = int(r[hdr.index("Instructions Executed")] or 0); sm = int(r[hdr.index("# Samples")] or 0)
    if ie >= mn: print(f"{r[0][-5:]} {ie:10d} {sm:6d} {r[1].strip()[:110]}")
