"""Writes profiles/ncu_traffic.json: measured DRAM traffic per launch of the kernel classes bench.py reports a roofline
for, from `ncu --set full` captures (dram__bytes_read.sum + dram__bytes_write.sum).

usage: python tools/ncu_traffic.py <attention .ncu-rep> [<decode .ncu-rep>]
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows_of(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = lambda k: next(i for i, h in enumerate(hdr) if h == k)
    out = []
    for d in data:
        def val(k):
            i = col(k)
            v = float(d[i].replace(",", ""))
            u = units[i].lower()
            return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        out.append((d[col("Kernel Name")], val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
                    float(d[col("gpu__time_duration.sum")].replace(",", "")), units[col("gpu__time_duration.sum")]))
    return out


def main():
    res = {}
    att = rows_of(sys.argv[1])
    bwd = [r for r in att if "relattn_bwd" in r[0]]
    fwd = [r for r in att if "relattn_fwd" in r[0]]
    if bwd:   # one launch of bench.py's attn_bwd class = one commu_relattn_bwd call = every kernel of that call together
        n_p1 = sum(1 for r in bwd if "bwd_p1" in r[0])          # materialised path: one pass-1 launch per call
        per_call = n_p1 if n_p1 else max(1, len(bwd) // 3)       # recompute path: three passes per call
        res["attn_bwd"] = {"bytes_per_launch": round(sum(r[1] for r in bwd) / per_call), "launches": len(bwd),
                           "kernels": {r[0].split("(")[0].split("::")[-1]: round(r[1]) for r in bwd},
                           "source": os.path.basename(sys.argv[1])}
    if fwd:
        res["attn_fwd"] = {"bytes_per_launch": round(sum(r[1] for r in fwd) / len(fwd)), "launches": len(fwd),
                           "source": os.path.basename(sys.argv[1])}
    if len(sys.argv) > 2:
        dec = [r for r in rows_of(sys.argv[2]) if "dec_attn" in r[0]]
        if dec:
            res["decode_attn"] = {"bytes_per_launch": round(sum(r[1] for r in dec) / len(dec)), "launches": len(dec),
                                  "source": os.path.basename(sys.argv[2])}
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    old = {}
    if os.path.exists(path):
        old = json.load(open(path))
    old.update(res)
    json.dump(old, open(path, "w"), indent=1)
    print(json.dumps(old, indent=1))


if __name__ == "__main__":
    main()
