#!/bin/bash
# round 2, last call: full GPU suite + smoke + default train bench line (no decode / baselines) on HEAD
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | tail -12 > gpurun_out/r02_gpu_tests.log; tail -3 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
DROPATT=0.1 timeout 300 python tools/time_attn.py 16 9 2>/dev/null | tail -1 > gpurun_out/r02_time_attn.json; cat gpurun_out/r02_time_attn.json
timeout 600 python bench.py --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_train_only_head.json 2>/dev/null
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02_bench_train_only_head.json')); print({k:j.get(k) for k in ("value","ms_per_step","kernel_time_ms_per_step","final_loss")})
PY
