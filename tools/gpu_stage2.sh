#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu 2>&1 | tail -40 > gpurun_out/tests_model.log; tail -6 gpurun_out/tests_model.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "relattn_fwd" 2>&1 | tail -40 > gpurun_out/tests_tc.log; tail -6 gpurun_out/tests_tc.log
if grep -q "failed" gpurun_out/tests_tc.log; then echo "tc forward FAILED - skipping tc bench"; else
COMMU_ATTN_FWD=tc timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err; tail -c 2500 gpurun_out/bench_tc.json; tail -3 gpurun_out/bench_tc.err
fi
