#!/bin/bash
# round 2, call I: pass 1 with two tiles of per-row constants in flight; attention tests; timing; bench (train only)
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -m gpu --maxfail=5 -k "relattn_bwd or dropout_fwd_bwd" 2>&1 | tail -3
DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 7 2>&1 | tail -1
timeout 900 python bench.py --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2i_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","clocks","final_loss")})
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/r2i_bench.err
