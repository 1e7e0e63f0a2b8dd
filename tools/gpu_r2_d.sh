#!/bin/bash
# round 2, call D: attention timing after the pass-1 L2 prefetch, decode parity at the bench shape, launch list of a train step
set +e
mkdir -p gpurun_out
DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -1
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_kernels_gpu.py -q -m gpu -k "bench_shape or relattn_bwd" 2>&1 | tail -5
cat gpurun_out/decode_parity_*.json
echo "--- launch list of one train step"
COMMU_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_launches_bench.csv python bench.py --steps 1 --warmup 2 --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2d_bench_under_ncu.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/r2d_launches_bench.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
ids = [int(r[0]) for r in rows[hi+1:] if r and r[0].isdigit()]
n = max(ids) + 1
print("launches", n)
# bench runs warmup 2 + 1 timed + 1 e2e step = 4 steps: take the third quarter
import subprocess
per = n // 4
print(subprocess.run(["python", "tools/launch_summary.py", "gpurun_out/r2d_launches_bench.csv", "gpurun_out/r2d_launch_shares.md", str(2*per), str(3*per)], capture_output=True, text=True).stdout)
PY
