#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_decode_gpu.py -q -m gpu -k "relattn_fwd or decode or sampler or generate or inference" 2>&1 | tail -25 > gpurun_out/tests_s3.log; tail -5 gpurun_out/tests_s3.log
COMMU_ATTN_FWD=tc timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/bench_tc.json'))
    print({k:j[k] for k in ("value","ms_per_step","kernel_time_ms_per_step")}); print(j.get("decode"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 gpurun_out/bench_tc.err
