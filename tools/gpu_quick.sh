#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_dropout_gpu.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py --no-decode --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json; j=json.load(open('gpurun_out/bench_q.json')); print(j['value'], j['ms_per_step'], j['kernel_time_ms_per_step'], j['gpu_launches'])"; tail -3 gpurun_out/bench_q.err
