#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_decode_gpu.py -q -m gpu -x 2>&1 | tail -25 > gpurun_out/tests_decode.log; tail -3 gpurun_out/tests_decode.log
COMMU_DECODE_PDL=1 timeout 300 python tools/decode_chain_probe.py 2>&1 | tail -1
for cfg in "8 2"; do
  set -- $cfg
  COMMU_DECODE_ATTN=$1 COMMU_DECODE_SPLITS=$2 COMMU_BENCH_FAST_PREFILL=1 timeout 300 python bench.py --decode-only 2>&1 | tail -1 | cut -c1-600 | sed "s/^/attn=$1 splits=$2 /"
done
