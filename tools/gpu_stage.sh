#!/bin/bash
# Runs the GPU validation stages on the B200 box; every stage logs into gpurun_out/ independently.
set +e
mkdir -p gpurun_out
echo "== tests ==" 
timeout 900 python -m pytest tests -q -m gpu -x --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | tail -40 > gpurun_out/tests_gpu.log
tail -15 gpurun_out/tests_gpu.log
echo "== smoke =="
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
echo "== bench =="
timeout 900 python bench.py --steps ${BENCH_STEPS:-4} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
