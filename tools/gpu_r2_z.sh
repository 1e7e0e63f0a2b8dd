#!/bin/bash
# round 2, call Z: compute-sanitizer memcheck over the model-level, dropout, decode and generation tests (small shapes)
set +e
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 7 \
  python -m pytest tests/test_model_gpu.py tests/test_dropout_gpu.py tests/test_decode_gpu.py tests/test_generate_gpu.py tests/test_checkpoint_gpu.py -q -m gpu -x \
  -k "not bench_shape and not 2048 and not perf" > gpurun_out/r02_memcheck_model.log 2>&1
echo "memcheck rc=$?"; tail -5 gpurun_out/r02_memcheck_model.log
