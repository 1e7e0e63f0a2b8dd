#!/bin/bash
# round 2, call R: per-shape GEMM timings (one-CTA vs CTA-pair kernel) and an ncu capture of the K = 512 shapes
set +e
mkdir -p gpurun_out
timeout 600 python tools/time_gemm_shapes.py 2>&1 | tail -14
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -c 18 -o gpurun_out/r2r_prof_gemm -f python tools/prof_gemm.py > gpurun_out/r2r_ncu.log 2>&1; tail -1 gpurun_out/r2r_ncu.log
