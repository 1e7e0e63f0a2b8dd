#!/bin/bash
# round 2, call G: fp32 decode engine (tiled linears, unrolled attention): decode tests + decode-only bench
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_generate_gpu.py -q -m gpu --maxfail=6 2>&1 | grep -v "^E    \+\|Warning\|warnings.warn" | tail -30
timeout 600 python bench.py --decode-only > gpurun_out/r2g_decode.json 2> gpurun_out/r2g_decode.err; cat gpurun_out/r2g_decode.json; tail -3 gpurun_out/r2g_decode.err
