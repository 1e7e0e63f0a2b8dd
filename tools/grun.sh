#!/bin/bash
# build + sanity-check locally, then run a command on the GPU box:  tools/grun.sh <timeout> '<command>' [--gpus N]
set -e
cd "$(dirname "$0")/.."
python commu-code_b200/build.py > /tmp/build.log 2>&1 || { tail -20 /tmp/build.log; echo "BUILD FAILED"; exit 1; }
test -f commu-code_b200/lib/libcommu_b200.so || { echo "no .so"; exit 1; }
python -m pytest tests/test_host_cpu.py tests/test_plumbing_dryrun.py -q -x > /tmp/cpu_tests.log 2>&1 || { tail -20 /tmp/cpu_tests.log; echo "CPU TESTS FAILED"; exit 1; }
exec /usr/local/graft/bin/gpurun $3 $4 --timeout "$1" -- "$2"
