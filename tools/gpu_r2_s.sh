#!/bin/bash
# round 2, call S: what paces the K = 512 GEMMs?  epilogue switched off piece by piece (COMMU_GEMM_EPI_DEBUG, measurement only)
set +e
for d in 0 1 2; do echo "== COMMU_GEMM_EPI_DEBUG=$d"; COMMU_GEMM_EPI_DEBUG=$d timeout 300 python tools/time_gemm_shapes.py 2>&1 | head -8; done
