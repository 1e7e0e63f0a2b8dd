#!/bin/bash
set +e
mkdir -p gpurun_out
COMMU_ATTN_BWD_DKV=tc timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "relattn_bwd" 2>&1 | tail -25 > gpurun_out/tests_s4.log; tail -8 gpurun_out/tests_s4.log
if grep -q "failed\|error" gpurun_out/tests_s4.log; then echo "dkv tc FAILED"; else
COMMU_ATTN_BWD_DKV=tc COMMU_ATTN_FWD=tc timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench_s4.json 2> gpurun_out/bench_s4.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_s4.json')); print({k:j[k] for k in ("value","ms_per_step","kernel_time_ms_per_step","final_loss")})
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_s4.err
fi
