#!/bin/bash
# round 2, call A: full GPU test suite (incl. the new bench-shape parity cases) + default bench line (baseline of the round)
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | tail -30 > gpurun_out/r2a_tests.log; tail -6 gpurun_out/r2a_tests.log
timeout 900 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2a_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","cpu_baseline","clocks","gpu_launches","reference_gpu","vs_reference_gpu")}); print(j.get("decode"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/r2a_bench.err
