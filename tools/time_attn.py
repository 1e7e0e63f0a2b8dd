"""Per-kernel timing of the tcgen05 attention kernels at the bench shape (CUDA events, not under a profiler).

usage: python tools/time_attn.py [B] [iters]   ->  one JSON line {kernel: ms, ...}
"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = M = 2048; H, Dh = 8, 64; K = T + M
dev = "cuda"; torch.manual_seed(0)
q = torch.randn(T, B, H * Dh, device=dev).bfloat16()
kv = torch.randn(K, B, 2 * H * Dh, device=dev).bfloat16()
r = torch.randn(K, H * Dh, device=dev).bfloat16()
u = torch.randn(H, Dh, device=dev) * 0.1; vb = torch.randn(H, Dh, device=dev) * 0.1
out = torch.empty(T, B, H * Dh, device=dev, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=dev)
qu = torch.empty_like(q); qv = torch.empty_like(q)
dout = torch.randn(T, B, H * Dh, device=dev).bfloat16()
delta = torch.zeros(B, H, T, device=dev); dq = torch.empty_like(q); dkv = torch.empty_like(kv)
dr = torch.zeros(K, H * Dh, device=dev); du = torch.zeros(H, Dh, device=dev); dvb = torch.zeros(H, Dh, device=dev)
sc = 1 / math.sqrt(Dh)
vv = kv[:, :, H * Dh:]; dvv = dkv[:, :, H * Dh:]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2


def fwd():
    nv.call("commu_relattn_fwd_tc", q, H * Dh, kv, vv, 2 * H * Dh, r, H * Dh, K, u, vb, None,
            T, M, B, H, 0, T, sc, out, H * Dh, lse, qu, qv)


def dq_():
    nv.call("commu_relattn_bwd_dq_tc", qu, qv, H * Dh, kv, vv, 2 * H * Dh, r, H * Dh, K, None,
            T, M, B, H, 0, T, sc, lse, dout, H * Dh, delta, dq, H * Dh, du, dvb)


def dkv_():
    nv.call("commu_relattn_bwd_dkv_tc", qu, qv, H * Dh, kv, vv, 2 * H * Dh, r, H * Dh, K, None,
            T, M, B, H, 0, T, sc, lse, dout, H * Dh, delta, dkv, dvv, 2 * H * Dh)


def dr_():
    nv.call("commu_relattn_bwd_dr_tc", qu, qv, H * Dh, kv, vv, 2 * H * Dh, r, H * Dh, K, None,
            T, M, B, H, 0, T, sc, lse, dout, H * Dh, delta, dr, du, dvb)


PD = float(os.environ.get("DROPATT", "0"))
nv.call("commu_relattn_set_dropout", PD, 12345678901234567)
res = {"B": B, "dropatt": PD}
fwd()
for name, fn in (("fwd", fwd), ("dq", dq_), ("dkv", dkv_), ("dr", dr_)):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(IT):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = round(sorted(ts)[len(ts) // 2], 4)
print(json.dumps(res))
