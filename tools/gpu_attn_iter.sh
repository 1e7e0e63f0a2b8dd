#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dropout_gpu.py -q -m gpu -x 2>&1 | tail -6 > gpurun_out/tests_attn.log; tail -4 gpurun_out/tests_attn.log
timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -1
DROPATT=0.1 timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -1
