"""ncu driver: tcgen05 attention forward (probabilities stored) + materialised backward at the BASELINE configs[1]
shape (B from argv, DROPATT env).  Kernels: relattn_fwd_tc_kernel, relattn_bwd_p1_kernel, relattn_bwd_band_kernel<0|1|2>."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = M = 2048; H, Dh = 8, 64; K = T + M
dev = "cuda"; torch.manual_seed(0)
q = torch.randn(T, B, H * Dh, device=dev).bfloat16()
kv = torch.randn(K, B, 2 * H * Dh, device=dev).bfloat16()
r = torch.randn(K, H * Dh, device=dev).bfloat16()
u = torch.randn(H, Dh, device=dev) * 0.1; vb = torch.randn(H, Dh, device=dev) * 0.1
out = torch.empty(T, B, H * Dh, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
qu = torch.empty_like(q); qv = torch.empty_like(q)
dout = torch.randn(T, B, H * Dh, device=dev).bfloat16()
delta = torch.empty(B, H, T, device=dev); dq = torch.empty_like(q); dkv = torch.empty_like(kv)
dr = torch.zeros(K, H * Dh, device=dev); du = torch.zeros(H, Dh, device=dev); dvb = torch.zeros(H, Dh, device=dev)
sc = 1 / math.sqrt(Dh)
p_bytes, mt_bytes, _, _ = nv.attn_sizes(T, M, B, H)
psv = torch.empty(p_bytes, dtype=torch.uint8, device=dev); mtv = torch.empty(mt_bytes // 4, device=dev)
ws = nv.attn_bwd_workspace(T, M, B, H, dev)
PD = float(os.environ.get("DROPATT", "0.1"))     # the bench default (reference attention_dropout 0.1)
nv.call("commu_relattn_set_dropout", PD, 0x1234567)
for it in range(2):
    nv.call("commu_relattn_fwd_tc", q, H * Dh, kv, kv[:, :, H * Dh:], 2 * H * Dh, r, H * Dh, K, u, vb, None,
            T, M, B, H, 0, T, sc, out, H * Dh, lse, qu, qv, psv, mtv)
    nv.call("commu_relattn_bwd", qu, qv, H * Dh, kv, kv[:, :, H * Dh:], 2 * H * Dh, r, H * Dh, K, None,
            T, M, B, H, 0, T, sc, out, H * Dh, lse, dout, H * Dh, delta, dq, H * Dh, dkv, dkv[:, :, H * Dh:],
            2 * H * Dh, dr, du, dvb, psv, mtv, ws, ws.numel())
torch.cuda.synchronize(); print("done")
