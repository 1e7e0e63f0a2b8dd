"""Times every GEMM shape of one training step at BASELINE configs[1] (B=16) with the one-CTA (128x256 tile) and the
CTA-pair (256x256, cta_group::2) kernels: CUDA events, median of 7 with an L2 flush between.  TFLOP/s per shape."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from commu import _native as nv
dev = "cuda"; torch.manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[3] * 1e3
R = 32768       # new rows per step (T * B)
KV = 65536      # key rows ((T + M) * B)
shapes = [     # name, m, n, k, a_mn, b_mn, split_k, out
    ("q_proj        ", R, 512, 512, 0, 0, 1, "bf16"),
    ("kv_proj       ", KV, 1024, 512, 0, 0, 1, "bf16"),
    ("o_net         ", R, 512, 512, 0, 0, 1, "f32"),
    ("ff1           ", R, 2048, 512, 0, 0, 1, "bf16"),
    ("ff2           ", R, 512, 2048, 0, 0, 1, "f32"),
    ("dgrad ff2     ", R, 2048, 512, 0, 1, 1, "bf16"),
    ("dgrad ff1     ", R, 512, 2048, 0, 1, 1, "f32"),
    ("dgrad kv      ", KV, 512, 1024, 0, 1, 1, "f32"),
    ("wgrad ff1 s16 ", 2048, 512, R, 1, 1, 16, "atomic"),
    ("wgrad ff2 s16 ", 512, 2048, R, 1, 1, 16, "atomic"),
    ("wgrad kv  s16 ", 1024, 512, KV, 1, 1, 16, "atomic"),
    ("wgrad q   s32 ", 512, 512, R, 1, 1, 32, "atomic"),
]
res = {}
for name, m, n, k, amn, bmn, sk, out in shapes:
    a = torch.randn((k, m) if amn else (m, k), device=dev).bfloat16()
    b = torch.randn((k, n) if bmn else (n, k), device=dev).bfloat16()
    kw = dict(m=m, n=n, k=k, a_mn=bool(amn), b_mn=bool(bmn), split_k=sk)
    if out == "bf16": kw["out_bf16"] = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    else: kw["out_f32"] = torch.zeros(m, n, device=dev); kw["f32_atomic"] = out == "atomic"
    row = {}
    for impl, tag in ((3, "one_cta"), (2, "pair")):
        try:
            us = t(lambda: nv.gemm(a, b, impl=impl, **kw))
            row[tag] = [round(us, 1), round(2.0 * m * n * k / us / 1e6, 0)]
        except Exception as e:
            row[tag] = str(e)[:60]
    res[name.strip()] = row
    print(name, row, flush=True)
print(json.dumps(res))
