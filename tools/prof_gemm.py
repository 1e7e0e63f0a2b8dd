"""Launches representative GEMM shapes for an ncu capture: python tools/prof_gemm.py [impl]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = "cuda"
for (m, n, k) in [(32768, 2048, 512), (32768, 512, 2048)]:
    a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
    out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        nv.gemm(a, b, m=m, n=n, k=k, out_bf16=out, impl=impl)
    torch.cuda.synchronize()
print("done")
