#!/bin/bash
# round 2, call O: forward with O accumulating in TMEM (lazy rescale): attention + model parity, timing, train bench
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py -q -m gpu --maxfail=5 2>&1 | grep -v "^E    \+\|Warning\|warnings.warn" | tail -6
DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 7 2>&1 | tail -1
ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 7 2>&1 | tail -1
timeout 900 python bench.py --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2o_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","clocks","final_loss")})
except Exception as e: print("bench parse failed", e)
PY
