"""Launches the FF1-style GEMM (bias + ReLU epilogue, bf16 out) for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
dev = "cuda"
m, n, k = 32768, 2048, 512
a = torch.randn(m, k, device=dev).bfloat16(); b = torch.randn(n, k, device=dev).bfloat16()
bias = torch.randn(n, device=dev)
out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
for _ in range(3):
    nv.gemm(a, b, m=m, n=n, k=k, bias=bias, relu=True, out_bf16=out)
torch.cuda.synchronize()
print("done")
