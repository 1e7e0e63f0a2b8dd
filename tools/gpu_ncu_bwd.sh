#!/bin/bash
# one ncu --set full capture of the three tcgen05 backward passes (second iteration) at B=4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:relattn_bwd_.*tc_kernel -s 3 -c 3 \
  -o gpurun_out/prof_bwd3 -f python tools/prof_bwd.py 4 > gpurun_out/ncu_bwd3.log 2>&1
tail -3 gpurun_out/ncu_bwd3.log
