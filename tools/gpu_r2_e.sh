#!/bin/bash
# round 2, call E: pass-1 ring restructure; full suite incl. generation / checkpoint tests; default bench; configs[4] bench at N=1
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -m gpu --maxfail=5 -k "relattn_bwd or dropout_fwd_bwd" 2>&1 | tail -5
DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -1
echo "--- full suite"
timeout 1800 python -m pytest tests -q -m gpu --maxfail=10 --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | grep -v "^E    \+\|Warning\|warnings.warn" | tail -60 > gpurun_out/r2e_tests.log; tail -25 gpurun_out/r2e_tests.log
echo "--- bench default"
timeout 1200 python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2e_bench.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","cpu_baseline","clocks","gpu_launches","vs_reference_gpu","final_loss")}); print(j.get("decode"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/r2e_bench.err
echo "--- bench c5 (configs[4]) N=1"
timeout 900 python bench.py --config c5 --steps 4 --warmup 3 > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r2e_bench_c5.json')); print({k:j.get(k) for k in ("metric","value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","step_mfu","clocks","final_loss")})
except Exception as e: print("bench c5 parse failed", e)
PY
tail -3 gpurun_out/r2e_bench_c5.err
