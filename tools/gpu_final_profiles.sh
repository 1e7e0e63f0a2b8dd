#!/bin/bash
# launch list of the bench command (shares, not absolutes): warm-up step + the timed step of a shortened run
set +e
mkdir -p gpurun_out
COMMU_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8600 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-decode > gpurun_out/bench_under_ncu.log 2>&1
tail -1 gpurun_out/bench_under_ncu.log | cut -c1-120; wc -l gpurun_out/launches_bench.csv
