"""Static Blackwell-opcode table of the built library (no GPU needed):
   python tools/sass_opcode_table.py [lib.so] > profiles/r02_sass_opcodes.md
Counts, per kernel of libcommu_b200.so, the SASS mnemonics that prove the tcgen05 / TMEM / TMA path
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG / UTMASTG / UTMAPF / UBLKCP,
legacy mma.sync -> HMMA)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "commu-code_b200", "lib", "libcommu_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
ops = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "UTMACMDFLUSH", "HMMA", "MUFU.EX2", "SYNCS", "REDG", "RED."]
cur, table = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        cur = table.setdefault(name, collections.Counter())
        continue
    if cur is None:
        continue
    for o in ops:
        if re.search(r"\b" + re.escape(o), line):
            cur[o.rstrip(".")] += 1
    if re.search(r"^\s*/\*[0-9a-f]{4,}\*/", line):
        cur["_n"] += 1
cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "HMMA", "MUFU.EX2", "REDG", "_n"]
print("# SASS opcode counts per kernel of libcommu_b200.so (cuobjdump -sass; static counts, tools/sass_opcode_table.py)\n")
print("| kernel | " + " | ".join(c if c != "_n" else "SASS instructions" for c in cols) + " |")
print("|---|" + "---:|" * len(cols))
for name, c in table.items():
    if not any(c[k] for k in cols[:8]):
        continue
    print("| `%s` | " % name[:70] + " | ".join(str(c[k]) if c[k] else "-" for k in cols) + " |")
print("\nKernels without any tensor-core / TMA opcode (SIMT elementwise, samplers, fp32 decode) are omitted.")
