"""Aggregate an ncu source page (cuda,sass) per CUDA source line:
   python tools/ncu_lines.py report.ncu-rep <kernel-name-regex> [top]
Every SASS row is attributed to the CUDA source line printed above it (needs -lineinfo at compile time)."""
import csv, subprocess, sys, io
rep, kre = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, cur_line, cur_src, hdr = None, None, "", None
agg = {}
first_fn = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) == 2 and r[0] == "Function Name":
        if first_fn is None: first_fn = r[1]
        elif r[1] != first_fn and False: break
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; iex = hdr.index("Instructions Executed"); ism = hdr.index("# Samples"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[0]:
        cur_line, cur_src = r[0], r[1].strip(); continue
    if not r[iex].isdigit(): continue
    key = (cur_file, cur_line)
    a = agg.setdefault(key, [0, 0, cur_src])
    a[0] += int(r[iex]); a[1] += int(r[ism])
te = sum(a[0] for a in agg.values()) or 1; ts = sum(a[1] for a in agg.values()) or 1
print("total warp instructions", te, "samples", ts)
for (f, l), a in sorted(agg.items(), key=lambda kv: -(kv[1][0] / te + kv[1][1] / ts))[:top]:
    print(f"{100*a[0]/te:5.1f}%exe {100*a[1]/ts:5.1f}%smp {f[:20]}:{l:>4} | {a[2][:100]}")
