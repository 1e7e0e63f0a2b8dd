"""Launches the K = 512 GEMM shapes of the training step (q projection bf16 out, o_net fp32 out + residual, FF1 bias +
ReLU bf16 out; one-CTA kernel, then the CTA-pair kernel) three times each, for an ncu capture."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
dev = "cuda"
m = 32768
a = torch.randn(m, 512, device=dev).bfloat16()
w5 = torch.randn(512, 512, device=dev).bfloat16(); w20 = torch.randn(2048, 512, device=dev).bfloat16()
bias = torch.randn(2048, device=dev); x = torch.randn(m, 512, device=dev)
ob = torch.empty(m, 512, device=dev, dtype=torch.bfloat16); of = torch.empty(m, 512, device=dev)
oh = torch.empty(m, 2048, device=dev, dtype=torch.bfloat16)
for impl in (3, 2):
    for _ in range(3): nv.gemm(a, w5, m=m, n=512, k=512, out_bf16=ob, impl=impl)
    for _ in range(3): nv.gemm(a, w5, m=m, n=512, k=512, add_f32=x, out_f32=of, impl=impl)
    for _ in range(3): nv.gemm(a, w20, m=m, n=2048, k=512, bias=bias, relu=True, out_bf16=oh, impl=impl)
torch.cuda.synchronize()
print("done")
