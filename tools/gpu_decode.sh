#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests_decode.log; tail -12 gpurun_out/tests_decode.log
for cfg in "0 1 0" "1 0 2" "1 1 2" "1 1 4" "1 1 6" "1 1 8"; do
  set -- $cfg
  COMMU_DECODE_FUSED=$1 COMMU_DECODE_PDL=$2 COMMU_DECODE_SPLITS=$3 timeout 300 python bench.py --decode-only 2>&1 | tail -1 | cut -c1-600 | sed "s/^/fused=$1 pdl=$2 splits=$3 /"
done
