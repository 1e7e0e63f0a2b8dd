"""Opcode mix of one kernel of an ncu report (SASS page): python tools/ncu_sass_mix.py rep kernel_index [top]"""
import csv, subprocess, sys, io, collections
rep, kid = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
k = -1; hdr = None; mix = collections.defaultdict(lambda: [0, 0]); name = ""
for r in rows:
    if r and r[0] == "Kernel Name": k += 1; name = r[1] if k == kid else name; continue
    if r and r[0] == "Address": hdr = r; continue
    if k != kid or hdr is None or len(r) < len(hdr): continue
    src = r[1].strip()
    toks = src.split()
    op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
    op = ".".join(op.split(".")[:3]).rstrip(";")
    ie = int(r[hdr.index("Instructions Executed")] or 0); sm = int(r[hdr.index("# Samples")] or 0)
    mix[op][0] += ie; mix[op][1] += sm
te = sum(v[0] for v in mix.values()); ts = sum(v[1] for v in mix.values())
print(name[:80]); print("total warp-inst", te, "samples", ts)
for op, (ie, sm) in sorted(mix.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*ie/te:5.1f}%exe {100*sm/max(ts,1):5.1f}%smp {ie:12d} {op}")
