#!/bin/bash
# round 2, call P: how much of the step is dropout (RNG in the GEMM epilogues, the attention forward, LayerNorm backward)?
set +e
mkdir -p gpurun_out
for d in 0.1 0.0; do
timeout 600 python bench.py --dropout $d --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r2p_bench_d$d.json 2>/dev/null
python - <<PY
import json
j=json.load(open('gpurun_out/r2p_bench_d$d.json')); print("dropout $d", {k:j.get(k) for k in ("value","ms_per_step","kernel_time_ms_per_step")})
PY
done
