"""Times the GEMM epilogue variants the training step uses (CUDA events, median of 7 with an L2 flush between)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
dev = "cuda"; torch.manual_seed(0)
M = 32768
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return round(sorted(ts)[3] * 1e3, 1)
def mk(m, n, k):
    return torch.randn(m, k, device=dev).bfloat16(), torch.randn(n, k, device=dev).bfloat16()
res = {}
a, b = mk(M, 512, 512); x = torch.randn(M, 512, device=dev); z = torch.empty(M, 512, device=dev); ob = torch.empty(M, 512, device=dev, dtype=torch.bfloat16)
res["q_bf16_512x512"] = t(lambda: nv.gemm(a, b, m=M, n=512, k=512, out_bf16=ob))
res["o_f32_res_512x512"] = t(lambda: nv.gemm(a, b, m=M, n=512, k=512, add_f32=x, out_f32=z))
res["o_f32_res_drop"] = t(lambda: nv.gemm(a, b, m=M, n=512, k=512, add_f32=x, out_f32=z, drop_p=0.1, drop_seed=5))
a1, b1 = mk(M, 2048, 512); bias1 = torch.randn(2048, device=dev); h = torch.empty(M, 2048, device=dev, dtype=torch.bfloat16)
res["ff1_plain_bf16"] = t(lambda: nv.gemm(a1, b1, m=M, n=2048, k=512, out_bf16=h))
res["ff1_bias_bf16"] = t(lambda: nv.gemm(a1, b1, m=M, n=2048, k=512, bias=bias1, out_bf16=h))
res["ff1_relu_bf16"] = t(lambda: nv.gemm(a1, b1, m=M, n=2048, k=512, relu=True, out_bf16=h))
res["ff1_bias_relu_bf16"] = t(lambda: nv.gemm(a1, b1, m=M, n=2048, k=512, bias=bias1, relu=True, out_bf16=h))
res["ff1_bias_relu_drop"] = t(lambda: nv.gemm(a1, b1, m=M, n=2048, k=512, bias=bias1, relu=True, out_bf16=h, drop_p=0.1, drop_seed=5))
a2, b2 = mk(M, 512, 2048); bias2 = torch.randn(512, device=dev)
res["ff2_bias_res_f32"] = t(lambda: nv.gemm(a2, b2, m=M, n=512, k=2048, bias=bias2, add_f32=x, out_f32=z))
res["ff2_bias_res_drop"] = t(lambda: nv.gemm(a2, b2, m=M, n=512, k=2048, bias=bias2, add_f32=x, out_f32=z, drop_p=0.1, drop_seed=5))
res["dgrad_ff2_relumask"] = t(lambda: nv.gemm(a, b1.t().contiguous()[:512].t().contiguous() if False else a1[:, :512].contiguous(), m=M, n=2048, k=512, relu_mask=h, out_bf16=h) if False else nv.gemm(a1, b1, m=M, n=2048, k=512, relu_mask=h, out_bf16=h))
w = torch.zeros(2048, 512, device=dev)
at = torch.randn(M, 2048, device=dev).bfloat16(); bt = torch.randn(M, 512, device=dev).bfloat16()
res["wgrad_2048x512_split16"] = t(lambda: nv.gemm(at, bt, m=2048, n=512, k=M, a_mn=True, b_mn=True, split_k=16, out_f32=w, f32_atomic=True))
print(json.dumps(res))
