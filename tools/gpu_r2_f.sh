#!/bin/bash
# round 2, call F (2 GPUs): data-parallel correctness check, then the 2-GPU bench lines (configs[1] and configs[4])
set +e
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 2 --check > gpurun_out/r2f_check.json 2> gpurun_out/r2f_check.err; echo "check rc=$?"; cat gpurun_out/r2f_check.json; tail -3 gpurun_out/r2f_check.err
timeout 900 $TR --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.err
timeout 900 $TR --master-port 29513 bench.py --gpus 2 --config c5 --steps 4 --warmup 3 > gpurun_out/r2f_bench_c5_n2.json 2> gpurun_out/r2f_bench_c5_n2.err
python - <<'PY'
import json
for f in ("r2f_bench_n2", "r2f_bench_c5_n2"):
    try:
        j=json.load(open('gpurun_out/%s.json' % f)); print(f, {k:j.get(k) for k in ("metric","value","n_gpus","ms_per_step","e2e","kernel_time_ms_per_step","clocks","final_loss")})
    except Exception as e: print(f, "parse failed", e)
PY
tail -q -n 2 gpurun_out/r2f_bench_n2.err gpurun_out/r2f_bench_c5_n2.err
