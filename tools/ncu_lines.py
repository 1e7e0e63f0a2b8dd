"""Aggregate an ncu source page (cuda,sass) per CUDA source line: python tools/ncu_lines.py rep kernel_index [top]"""
import csv, subprocess, sys, io
rep, kid = sys.argv[1], int(sys.argv[2]); top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file = None; hdr = None; out = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    g = lambda k: int(r[hdr.index(k)]) if r[hdr.index(k)].lstrip("-").isdigit() else 0
    st = {k: g(k) for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_sleep", "stall_wait", "stall_math", "stall_not_selected", "stall_selected", "stall_branch_resolving", "stall_no_inst")}
    out.append((g("Instructions Executed"), g("# Samples"), cur_file, r[0], r[1].strip()[:100], st))
te = sum(o[0] for o in out); ts = sum(o[1] for o in out)
print("total inst", te, "samples", ts)
for o in sorted(out, key=lambda o: -(o[0] / te + o[1] / ts))[:top]:
    big = ",".join(f"{k[6:]}={v}" for k, v in sorted(o[5].items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{100*o[0]/te:5.1f}%exe {100*o[1]/ts:5.1f}%smp {o[2][:18]}:{o[3]:>4} | {o[4][:80]} | {big}")
