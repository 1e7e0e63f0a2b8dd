#!/bin/bash
# round 2, call H: ncu --set full of pass 1 (3-deep ring) with source counters
set +e
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relattn_bwd_p1" -s 1 -c 1 -o gpurun_out/r2h_prof_p1 -f python tools/prof_bwd.py 16 > gpurun_out/r2h_ncu.log 2>&1; tail -2 gpurun_out/r2h_ncu.log
