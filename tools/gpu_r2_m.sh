#!/bin/bash
# round 2, call M: launch list of the fp32 decode step
set +e
mkdir -p gpurun_out
COMMU_BENCH_FAST_PREFILL=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches_decode.csv python bench.py --decode-only > gpurun_out/r2m_decode_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/r2m_launches_decode.csv gpurun_out/r2m_launch_shares_decode.md 150 400
