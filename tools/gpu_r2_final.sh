#!/bin/bash
# round 2, evidence run on one B200: GPU test suite, smoke, reference arm, default bench line, launch list + ncu summary
set +e
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu --deselect tests/test_gemm_gpu.py::test_gemm_perf 2>&1 | tail -12 > gpurun_out/r02_gpu_tests.log; tail -4 gpurun_out/r02_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; cut -c1-400 gpurun_out/r02_bench_reference_arm.json
timeout 1500 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err
python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r02_bench_default_final.json')); print({k:j.get(k) for k in ("value","ms_per_step","e2e","kernel_time_ms_per_step","roofline","cpu_baseline","clocks","gpu_launches","vs_reference_gpu","final_loss")}); print(j.get("decode"))
except Exception as e: print("bench parse failed", e)
PY
tail -3 gpurun_out/r02_bench_default_final.err
timeout 600 python bench.py --batch-chunk 4 --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_batch_chunk4.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_batch_chunk4.json
timeout 900 python bench.py --config c5 --steps 4 --warmup 3 > gpurun_out/r02_bench_c5_n1.json 2>/dev/null; cut -c1-300 gpurun_out/r02_bench_c5_n1.json
timeout 300 python tools/time_ln.py 2>/dev/null | tail -1 > gpurun_out/r02_time_ln.json
COMMU_BENCH_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_ncu_launches_bench.csv python bench.py --steps 1 --warmup 2 --no-decode --no-cpu-baseline --no-reference-gpu > gpurun_out/r02_bench_under_ncu.log 2>&1
python - <<'PY'
import csv, subprocess
rows = list(csv.reader(open('gpurun_out/r02_ncu_launches_bench.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
n = max(int(r[0]) for r in rows[hi+1:] if r and r[0].isdigit()) + 1
per = n // 4
print(subprocess.run(["python", "tools/launch_summary.py", "gpurun_out/r02_ncu_launches_bench.csv", "gpurun_out/r02_launch_shares_bench.md", str(2*per), str(3*per)], capture_output=True, text=True).stdout)
PY
# forward + the three kernels of one materialised backward (pass 1, merged dq_A/dq_C, dR): second iteration of prof_bwd.py
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"relattn_bwd_p1|relattn_bwd_band|relattn_fwd_tc" -s 4 -c 4 -o gpurun_out/r02_prof_attn_mat -f python tools/prof_bwd.py 16 > gpurun_out/r02_ncu.log 2>&1; tail -1 gpurun_out/r02_ncu.log
# the K = 512 GEMM shapes (one-CTA and CTA-pair kernels, whole-tile epilogues)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tcgen05" -c 18 -o gpurun_out/r02_prof_gemm -f python tools/prof_gemm.py > gpurun_out/r02_ncu_gemm.log 2>&1; tail -1 gpurun_out/r02_ncu_gemm.log
timeout 300 python tools/time_gemm_shapes.py 2>/dev/null | tail -1 > gpurun_out/r02_gemm_shapes.json
DROPATT=0.1 timeout 300 python tools/time_attn.py 16 5 2>/dev/null | tail -1 > gpurun_out/r02_time_attn.json; cat gpurun_out/r02_time_attn.json
