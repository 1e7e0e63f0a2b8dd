#!/bin/bash
# round 2, call X: software-pipelined LayerNorm backward - parity, timing, bench line
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_dropout_gpu.py -q -m gpu -x -k "not relattn" 2>&1 | tail -3
timeout 300 python tools/time_ln.py 2>&1 | tail -1
timeout 600 python bench.py --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2x_bench.json 2>/dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2x_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","kernel_time_ms_per_step","final_loss")})
PY
