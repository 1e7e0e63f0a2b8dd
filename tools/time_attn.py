"""Device timing (CUDA events, L2 flushed between runs) of the attention kernels at the BASELINE configs[1] shape:
forward with / without the stored probabilities, the materialised backward (commu_relattn_bwd with p_save) and the
three recompute passes.  usage: time_attn.py [B] [iters]   (DROPATT=0.1 for attention dropout)"""
import json, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
IT = int(sys.argv[2]) if len(sys.argv) > 2 else 5
T = int(os.environ.get("ATT_T", "2048")); M = int(os.environ.get("ATT_M", str(T))); H = int(os.environ.get("ATT_H", "8"))
Dh = 64; K = T + M
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
vv = kv[:, :, H * Dh:]; dvv = dkv[:, :, H * Dh:]
p_bytes, mt_bytes, ws_bytes, _ = nv.attn_sizes(T, M, B, H)
psv = torch.empty(p_bytes, dtype=torch.uint8, device=dev); mtv = torch.empty(mt_bytes // 4, device=dev)
ws = nv.attn_bwd_workspace(T, M, B, H, dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2


def fwd(store=True):
    nv.call("commu_relattn_fwd_tc", q, H * Dh, kv, vv, 2 * H * Dh, r, H * Dh, K, u, vb, None,
            T, M, B, H, 0, T, sc, out, H * Dh, lse, qu, qv, psv if store else None, mtv if store else None)


def bwd(mat=True):
    nv.call("commu_relattn_bwd", qu, qv, H * Dh, kv, vv, 2 * H * Dh, r, H * Dh, K, None,
            T, M, B, H, 0, T, sc, out, H * Dh, lse, dout, H * Dh, delta, dq, H * Dh, dkv, dvv,
            2 * H * Dh, dr, du, dvb, psv if mat else None, mtv if mat else None, ws if mat else None,
            ws.numel() if mat else 0)


PD = float(os.environ.get("DROPATT", "0"))
nv.call("commu_relattn_set_dropout", PD, 12345678901234567)
res = {"B": B, "T": T, "M": M, "H": H, "dropatt": PD, "p_save_GB": round(p_bytes / 2 ** 30, 2), "ws_GB": round(ws_bytes / 2 ** 30, 2)}
fwd()
cases = [("fwd_store", lambda: fwd(True)), ("fwd_plain", lambda: fwd(False)), ("bwd_mat", lambda: bwd(True))]
if os.environ.get("ATT_LEGACY", "1") == "1":
    cases.append(("bwd_recompute", lambda: bwd(False)))
for name, fn in cases:
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(IT):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res[name] = round(sorted(ts)[len(ts) // 2], 4)
fwd(True)
print(json.dumps(res))
