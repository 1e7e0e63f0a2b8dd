#!/bin/bash
# round 2, 8 GPUs: configs[2] (12L d512, data parallel on 8 x B200) and configs[4] (24L d1024 T = M = 4096) bench lines + the DP check
set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 bench.py --gpus 8 --check > gpurun_out/r02_dp_check_n8.json 2> gpurun_out/r02_dp_check_n8.err; echo "check rc=$?"; grep '^{' gpurun_out/r02_dp_check_n8.json | cut -c1-600
timeout 900 $TR --master-port 29522 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
timeout 900 $TR --master-port 29523 bench.py --gpus 8 --config c5 --steps 4 --warmup 3 > gpurun_out/r02_bench_c5_n8.json 2> gpurun_out/r02_bench_c5_n8.err
python - <<'PY'
import json
for f in ("r02_bench_n8", "r02_bench_c5_n8"):
    try:
        for line in open('gpurun_out/%s.json' % f):
            if line.startswith('{'):
                j=json.loads(line); print(f, {k:j.get(k) for k in ("metric","value","n_gpus","ms_per_step","e2e","kernel_time_ms_per_step","clocks","final_loss")})
    except Exception as e: print(f, "parse failed", e)
PY
