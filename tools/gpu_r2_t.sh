#!/bin/bash
# round 2, call T: whole-tile GEMM epilogues - parity, per-shape timings, bench line
set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gemm_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout 300 python tools/time_gemm_shapes.py 2>&1 | head -12
timeout 300 python tools/time_gemm_variants.py 2>&1 | tail -1
timeout 600 python bench.py --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2t_bench.json 2>/dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2t_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","kernel_time_ms_per_step","final_loss")})
PY
