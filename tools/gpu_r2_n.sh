#!/bin/bash
# round 2, call N: ncu --set full of the fp32 decode kernels (attention + tiled linears) inside a token step
set +e
mkdir -p gpurun_out
COMMU_BENCH_FAST_PREFILL=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"decode_attn_kernel|decode_linear_tiled" -s 1700 -c 6 -o gpurun_out/r2n_prof_decode_fp32 -f python bench.py --decode-only > gpurun_out/r2n_ncu.log 2>&1; tail -2 gpurun_out/r2n_ncu.log
