#!/bin/bash
# round 2: ncu launch list of the timed train step + --set full capture of the attention kernels at HEAD
set +e
mkdir -p gpurun_out
COMMU_BENCH_PROFILE=1 timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 2 --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, subprocess
rows = list(csv.reader(open('gpurun_out/r02_ncu_launches_bench.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
n = max(int(r[0]) for r in rows[hi+1:] if r and r[0].isdigit()) + 1
per = n // 4
print(subprocess.run(["python", "tools/launch_summary.py", "gpurun_out/r02_ncu_launches_bench.csv", "gpurun_out/r02_launch_shares_bench.md", str(2*per), str(3*per)], capture_output=True, text=True).stdout)
PY
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"relattn_bwd_p1|relattn_bwd_band|relattn_fwd_tc" -s 4 -c 4 -o gpurun_out/r02_prof_attn_mat -f python tools/prof_bwd.py 16 > gpurun_out/r02_ncu.log 2>&1; tail -1 gpurun_out/r02_ncu.log
