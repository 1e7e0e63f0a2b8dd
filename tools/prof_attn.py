"""Profiling driver: runs the attention forward/backward and one GEMM at BASELINE configs[1] shapes
(T = M = 2048, H = 8, Dh = 64) for ncu captures.  `python tools/prof_attn.py [B] [reps]`"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T = M = 2048
H, Dh = 8, 64
K = T + M
dev = "cuda"
torch.manual_seed(0)
q = torch.randn(T, B, H * Dh, device=dev).bfloat16()
kv = torch.randn(K, B, 2 * H * Dh, device=dev).bfloat16()
r = torch.randn(K, H * Dh, device=dev).bfloat16()
u = torch.randn(H, Dh, device=dev) * 0.1
vb = torch.randn(H, Dh, device=dev) * 0.1
out = torch.empty(T, B, H * Dh, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device=dev)
qu = torch.empty_like(q)
qv = torch.empty_like(q)
dout = torch.randn(T, B, H * Dh, device=dev).bfloat16()
delta = torch.empty(B, H, T, device=dev)
dq = torch.empty_like(q)
dkv = torch.empty_like(kv)
dr = torch.zeros(K, H * Dh, device=dev)
du = torch.zeros(H, Dh, device=dev)
dvb = torch.zeros(H, Dh, device=dev)
scale = 1 / math.sqrt(Dh)
x = torch.randn(T * B, 512, device=dev).bfloat16()
w = torch.randn(2048, 512, device=dev).bfloat16()
y = torch.empty(T * B, 2048, device=dev, dtype=torch.bfloat16)
for it in range(reps):
    nv.call("commu_relattn_fwd", q, H * Dh, kv, kv[:, :, H * Dh:], 2 * H * Dh, r, H * Dh, K, u, vb, None,
            T, M, B, H, 0, T, scale, out, H * Dh, lse, qu, qv)
    nv.call("commu_relattn_bwd", qu, qv, H * Dh, kv, kv[:, :, H * Dh:], 2 * H * Dh, r, H * Dh, K, None,
            T, M, B, H, 0, T, scale, out, H * Dh, lse, dout, H * Dh, delta, dq, H * Dh, dkv, dkv[:, :, H * Dh:],
            2 * H * Dh, dr, du, dvb)
    nv.gemm(x, w, m=T * B, n=2048, k=512, out_bf16=y)
torch.cuda.synchronize()
print("done")
