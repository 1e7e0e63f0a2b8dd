#!/bin/bash
# round 2, call B: materialised attention backward - kernel parity first, then timing, then the whole suite, then bench
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -m gpu --maxfail=8 -k "relattn" 2>&1 | tail -60 > gpurun_out/r2b_attn_tests.log; tail -25 gpurun_out/r2b_attn_tests.log
echo "--- timing"
timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -3
DROPATT=0.1 timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -3
echo "--- full suite"
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | tail -40 > gpurun_out/r2b_tests.log; tail -15 gpurun_out/r2b_tests.log
echo "--- bench"
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2b_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","cpu_baseline","clocks","gpu_launches","reference_gpu","vs_reference_gpu","final_loss")}); print(j.get("decode"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/r2b_bench.err
