#!/bin/bash
# round 2, call Q: merged dq_A / dq_C launch (L2 reuse of the dS rows) - parity, then timing against the two-launch form
set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "relattn or dropout or forward_backward or train" 2>&1 | tail -5
for m in 1 0; do
echo "== COMMU_ATTN_BAND_MERGE=$m"
COMMU_ATTN_BAND_MERGE=$m DROPATT=0.1 timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -4
done
COMMU_BENCH_PROFILE=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"relattn_bwd" -c 12 --csv --log-file gpurun_out/r2q_ncu_bwd.csv python tools/prof_bwd.py 16 > gpurun_out/r2q_ncu.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2q_ncu_bwd.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
h=rows[hi]
for r in rows[hi+1:]:
    if len(r)<len(h): continue
    print(r[h.index("Kernel Name")][:50], r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
PY
timeout 600 python bench.py --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2q_bench.json 2>/dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2q_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","kernel_time_ms_per_step","final_loss")})
PY
