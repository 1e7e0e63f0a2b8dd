"""Times CUDA-graph chains of each fused decode linear kernel alone (how much of a token step is per-kernel latency)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from types import SimpleNamespace as NS
from commu import _native as nv
from commu.model.model import MemTransformerLM
from commu.engine.decode import DecodeEngine

class V:
    def __len__(self): return 729
cfg = NS(MODEL=NS(num_layers=2, num_heads=8, units=512, inner_size=2048, dropout=0.0, attention_dropout=0.0, same_length=True, clamp_len=-1),
         TRAIN=NS(tgt_length=1, mem_length=2048))
m = MemTransformerLM(cfg, V()).cuda().eval()
eng = DecodeEngine(m, batch=64, mem_len=2048, same_length=True, precision="bf16")
tok = torch.randint(1, 700, (64,), device="cuda")
eng.fargs[0][0].tokens = tok.data_ptr()
eng.fargs[0][0].slot = 0
eng.fargs[1][0].slot = 0
names = ["qkv_embed", "o", "ff1", "ff2"]
res = {}
def chain(fn, n=48):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 / n * 1e3
for i, nm in enumerate(names):
    a = eng.fargs[0][i]
    res[nm] = round(chain(lambda a=a: nv.dec_linear(a)), 2)
res["qkv_ln"] = round(chain(lambda: nv.dec_linear(eng.fargs[1][0])), 2)
res["logits"] = round(chain(lambda: nv.dec_linear(eng.a_logits)), 2)
lg = eng.ws["logits"]; cur = torch.zeros(64, dtype=torch.int64, device="cuda")
res["sampler"] = round(chain(lambda: nv.call("commu_sample", lg, eng.V, 64, eng.V, 0.95, 0, 0.9, None, 1, 0, cur, None, eng.V, None)), 2)
dstate = torch.tensor([5, 2048, 2048, 0], dtype=torch.int32, device="cuda")
res["advance"] = round(chain(lambda: nv.call("commu_decode_advance", dstate, eng.C, eng.mem_len, 0)), 2)
res["layer_linears"] = round(chain(lambda: [nv.dec_linear(x) for x in eng.fargs[1]], n=12), 2)
print(json.dumps({"pdl": eng.pdl, "us_per_launch": res}))
