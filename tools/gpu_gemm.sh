#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x -k "not perf" 2>&1 | tail -15 > gpurun_out/tests_gemm.log; tail -4 gpurun_out/tests_gemm.log
COMMU_GEMM_2CTA=1 timeout 300 python -m pytest tests/test_gemm_gpu.py::test_gemm_perf -q -m gpu -s 2>&1 | grep -E "^\[|passed|failed" | cut -c1-1800
COMMU_GEMM_2CTA=0 timeout 300 python -m pytest tests/test_gemm_gpu.py::test_gemm_perf -q -m gpu -s 2>&1 | grep -E "^\[|passed|failed" | cut -c1-1800
