#!/bin/bash
timeout 300 python tools/time_gemm_variants.py 2>&1 | tail -1
