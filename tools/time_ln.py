"""CUDA-event timing of the LayerNorm kernels at the bench shape ([32768, 512] fp32, L2 flushed between runs)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
dev = "cuda"; torch.manual_seed(0)
n, d = 32768, 512
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(9):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return round(sorted(ts)[4] * 1e3, 1)
z = torch.randn(n, d, device=dev); dy = torch.randn(n, d, device=dev)
gam = torch.randn(d, device=dev); bet = torch.randn(d, device=dev)
y = torch.empty(n, d, device=dev); yb = torch.empty(n, d, device=dev, dtype=torch.bfloat16)
mean = torch.empty(n, device=dev); rstd = torch.empty(n, device=dev)
dzf = torch.empty(n, d, device=dev); dzb = torch.empty(n, d, device=dev, dtype=torch.bfloat16)
dg = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev)
res = {}
res["ln_fwd_us"] = t(lambda: nv.call("commu_layernorm_fwd", z, d, gam, bet, d, d, 1e-5, n, y, d, yb, d, mean, rstd))
res["ln_bwd_us"] = t(lambda: nv.call("commu_layernorm_bwd", dy, d, z, d, mean, rstd, gam, d, d, n, dzf, d, dzb, d, dg, db, 0.0, 0))
res["ln_bwd_drop_us"] = t(lambda: nv.call("commu_layernorm_bwd", dy, d, z, d, mean, rstd, gam, d, d, n, dzf, d, dzb, d, dg, db, 0.1, 1234))
res["bytes_bwd_MB"] = round((3 * n * d * 4 + n * d * 2) / 1e6, 1)
print(json.dumps(res))
