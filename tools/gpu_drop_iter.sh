#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropout_gpu.py tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests_drop.log; tail -12 gpurun_out/tests_drop.log
timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -1
DROPATT=0.1 timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -1
