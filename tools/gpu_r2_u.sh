#!/bin/bash
# round 2, call U: ncu capture of the K = 512 GEMM shapes with the whole-tile epilogues
set +e
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -c 18 -o gpurun_out/r2u_prof_gemm -f python tools/prof_gemm.py > gpurun_out/r2u_ncu.log 2>&1; tail -1 gpurun_out/r2u_ncu.log
