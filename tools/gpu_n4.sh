#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; cut -c1-300 gpurun_out/bench_n4.json; tail -3 gpurun_out/bench_n4.err
