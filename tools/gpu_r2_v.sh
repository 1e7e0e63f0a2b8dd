#!/bin/bash
# round 2, call V: attention kernel change - parity, timing, bench line
set +e
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py tests/test_model_gpu.py -q -m gpu -x -k "relattn or dropout or forward_backward or train" 2>&1 | tail -3
DROPATT=0.1 timeout 300 python tools/time_attn.py 16 9 2>&1 | tail -1
