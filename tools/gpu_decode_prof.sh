#!/bin/bash
# decode step: tests, launch list (per-kernel durations of one token step) and full captures of the fused kernels
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests_decode.log; tail -4 gpurun_out/tests_decode.log
export COMMU_BENCH_FAST_PREFILL=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"dec_|sampler|advance" --launch-skip 400 -c 140 --csv --log-file gpurun_out/launches_decode.csv python bench.py --decode-only > gpurun_out/launches_decode.log 2>&1; tail -2 gpurun_out/launches_decode.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dec_attn_split|dec_linear" --launch-skip 330 -c 6 -o gpurun_out/prof_decode -f python bench.py --decode-only > gpurun_out/ncu_decode.log 2>&1; tail -2 gpurun_out/ncu_decode.log | cut -c1-300
ls -la gpurun_out/*.ncu-rep
