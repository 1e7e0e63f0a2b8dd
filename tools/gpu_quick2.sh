#!/bin/bash
# usage: gpu_quick2.sh "<env assignments>" "<pytest -k expr>" "<bench args>"
set +e
mkdir -p gpurun_out
export $1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_decode_gpu.py tests/test_model_gpu.py -q -m gpu -k "$2" 2>&1 | tail -25 > gpurun_out/tests_q.log; tail -6 gpurun_out/tests_q.log
if grep -q "failed\|error" gpurun_out/tests_q.log 2>/dev/null; then echo "TESTS FAILED - skipping bench"; exit 0; fi
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline $3 > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_q.json')); print({k:j[k] for k in ("value","ms_per_step","kernel_time_ms_per_step","final_loss")}); print(j.get("decode"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_q.err
