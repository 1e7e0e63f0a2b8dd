#!/bin/bash
# attention-kernel iteration: bwd/fwd parity tests + per-kernel timing (+ optional ncu of the bwd kernels)
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "relattn" 2>&1 | tail -15 > gpurun_out/tests_attn.log; tail -8 gpurun_out/tests_attn.log
timeout 300 python tools/time_attn.py 16 5 2>&1 | tail -3 | tee gpurun_out/time_attn.json
if [ "$1" = "ncu" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relattn.*_tc" -c 4 -s 4 -o gpurun_out/prof_attn_iter -f python tools/prof_bwd.py 16 > gpurun_out/ncu_attn_iter.log 2>&1; tail -2 gpurun_out/ncu_attn_iter.log
fi
