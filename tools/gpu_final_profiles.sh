#!/bin/bash
# launch list of the bench command (shares, not absolutes) + full captures of the attention kernels at the bench shape
set +e
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode > gpurun_out/bench_under_ncu.log 2>&1
tail -2 gpurun_out/bench_under_ncu.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relattn.*_tc|attn_delta" -c 5 -s 5 -o gpurun_out/prof_attn_b16 -f python tools/prof_bwd.py 16 > gpurun_out/ncu_attn_b16.log 2>&1; tail -2 gpurun_out/ncu_attn_b16.log
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_bench.csv
