#!/bin/bash
# round 2, call M: launch list of the fp32 decode step (skips engine preparation; the token steps are graph replays)
set +e
mkdir -p gpurun_out
COMMU_BENCH_FAST_PREFILL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1400 -c 500 --csv --log-file gpurun_out/r2m_launches_decode.csv python bench.py --decode-only > gpurun_out/r2m_decode_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2m_launches_decode.csv gpurun_out/r2m_launch_shares_decode.md
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2m_launches_decode.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
for r in data[300:345]:
    if len(r) > vi: print(r[0], r[ki].split('(')[0].replace('void ','').replace('<unnamed>::','')[:40], r[gi], r[vi])
PY
