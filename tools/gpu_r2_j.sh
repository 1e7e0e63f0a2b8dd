#!/bin/bash
# round 2, call J: 4-D TMA boxes for the dS tile (pass 1 stores, dq_A loads); fused qkv epilogue of the fp32 decode engine
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -m gpu --maxfail=5 -k "relattn_bwd or dropout_fwd_bwd" 2>&1 | tail -3
DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 7 2>&1 | tail -1
COMMU_ATTN_DS4D=0 DROPATT=0.1 ATT_LEGACY=0 timeout 300 python tools/time_attn.py 16 7 2>&1 | tail -1
timeout 900 python -m pytest tests/test_decode_gpu.py tests/test_generate_gpu.py tests/test_model_gpu.py -q -m gpu --maxfail=6 2>&1 | grep -v "^E    \+\|Warning\|warnings.warn" | tail -8
timeout 600 python bench.py --decode-only > gpurun_out/r2j_decode.json 2> gpurun_out/r2j_decode.err; cut -c1-700 gpurun_out/r2j_decode.json; tail -3 gpurun_out/r2j_decode.err
