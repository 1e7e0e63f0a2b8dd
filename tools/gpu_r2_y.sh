#!/bin/bash
# round 2, call Y: compute-sanitizer memcheck over the small-shape attention (materialised backward, tcgen05 forward),
# GEMM-epilogue and LayerNorm tests - out-of-bounds / misaligned accesses of the kernels changed this round
set +e
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 \
  python -m pytest tests/test_kernels_gpu.py tests/test_gemm_gpu.py -q -m gpu -x \
  -k "(relattn_bwd and mat and not 2048 and not 4096) or (relattn_fwd and _tc and not 2048 and not 4096 and not large) or gemm_epilogue or elementwise or layernorm" \
  > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -c "Invalid\|misaligned\|out of bounds" gpurun_out/r02_memcheck.log; tail -6 gpurun_out/r02_memcheck.log
