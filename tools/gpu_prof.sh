#!/bin/bash
# tests (model only) + ncu captures of the attention / GEMM kernels
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu 2>&1 | tail -30 > gpurun_out/tests_model.log; tail -8 gpurun_out/tests_model.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relattn|gemm_tcgen05" -c 6 -o gpurun_out/prof_attn -f python tools/prof_attn.py 2 1 > gpurun_out/ncu_attn.log 2>&1; tail -3 gpurun_out/ncu_attn.log
ls -la gpurun_out/*.ncu-rep
