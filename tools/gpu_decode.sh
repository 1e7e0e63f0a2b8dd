#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests_decode.log; tail -3 gpurun_out/tests_decode.log
for sk in attn none; do
for pdl in 1 0; do
  COMMU_DECODE_SKIP=$sk COMMU_DECODE_PDL=$pdl COMMU_BENCH_FAST_PREFILL=1 timeout 300 python bench.py --decode-only 2>&1 | tail -1 | cut -c1-600 | sed "s/^/attn=$sk splits=$pdl /"
done
done
