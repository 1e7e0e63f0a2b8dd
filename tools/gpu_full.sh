#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | tail -30 > gpurun_out/tests_full.log; tail -4 gpurun_out/tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json; tail -2 gpurun_out/bench_ref.err
timeout 1200 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_default.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","cpu_baseline","clocks","gpu_launches")}); print(j.get("decode"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_default.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dec_attn_tma|dec_linear" -s 12 -c 6 -o gpurun_out/prof_decode_s3 -f python tools/prof_decode.py > gpurun_out/ncu_decode.log 2>&1; tail -2 gpurun_out/ncu_decode.log
