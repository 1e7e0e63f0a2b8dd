#!/bin/bash
# round 2, call L: fp32 decode (unrolled split-K finalize, key-split attention); pass-1 loop restored
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_generate_gpu.py -q -m gpu --maxfail=6 2>&1 | grep -v "^E    \+\|Warning\|warnings.warn" | tail -6
timeout 600 python bench.py --decode-only > gpurun_out/r2l_decode.json 2> gpurun_out/r2l_decode.err; cut -c1-700 gpurun_out/r2l_decode.json; tail -3 gpurun_out/r2l_decode.err
DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 7 2>&1 | tail -1
