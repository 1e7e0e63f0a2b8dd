#!/bin/bash
# decode step: launch list (per-kernel durations of one token step)
set +e
mkdir -p gpurun_out
export COMMU_BENCH_FAST_PREFILL=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dec_|sampler|advance" --launch-skip 400 -c 140 --csv --log-file gpurun_out/launches_decode.csv python bench.py --decode-only > gpurun_out/launches_decode.log 2>&1; tail -1 gpurun_out/launches_decode.log | cut -c1-200
