#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropout_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/tests_drop.log; tail -8 gpurun_out/tests_drop.log
timeout 600 python bench.py --no-decode --no-cpu-baseline > gpurun_out/bench_drop01.json 2> gpurun_out/bench_drop01.err; python -c "
import json; j=json.load(open('gpurun_out/bench_drop01.json')); print(j['value'], j['ms_per_step'], j['kernel_time_ms_per_step'], j['gpu_launches'], j['config']['workload'])"; tail -3 gpurun_out/bench_drop01.err
