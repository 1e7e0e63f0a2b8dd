#!/bin/bash
# round 2, call C: re-run the failed tests with detail, per-kernel launch list + ncu --set full of the materialised backward
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py tests/test_model_gpu.py tests/test_dropout_gpu.py -q -m gpu -k "bench_shape or train_steps_golden or dropout_fwd_bwd" 2>&1 | grep -v "^E    \+\|Warning\|warnings" | tail -60 > gpurun_out/r2c_tests.log; tail -40 gpurun_out/r2c_tests.log
echo "--- launch list (B=16, dropout 0.1)"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"relattn|rrev|delta" --csv --log-file gpurun_out/r2c_launches_attn.csv python tools/prof_bwd.py 16 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/r2c_launches_attn.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; ki, mi, vi = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    k = r[ki][:60]; agg.setdefault(k, collections.defaultdict(list))[r[mi]].append(float(r[vi].replace(',', '')))
for k, d in agg.items():
    print(k, {m: round(sum(v[len(v)//2:]) / max(1, len(v) - len(v)//2) / (1e6 if 'time' in m else 1e6), 3) for m, v in d.items()}, '(ms | MB, second half of launches)')
PY
echo "--- ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relattn_bwd_p1|relattn_bwd_band|relattn_fwd_tc" -s 5 -c 5 -o gpurun_out/r2c_prof_attn_mat -f python tools/prof_bwd.py 16 > gpurun_out/r2c_ncu.log 2>&1; tail -2 gpurun_out/r2c_ncu.log
