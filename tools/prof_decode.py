"""ncu driver for the fused decode kernels at the BASELINE configs[3] shape (B 64, d 512, H 8, mem 2048) with only two
layers (the per-launch work is identical; ncu's replay has 0.5 GB instead of 3.3 GB of cache to save / restore)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))
import torch
from types import SimpleNamespace as NS
from commu.model.model import MemTransformerLM
from commu.engine.decode import DecodeEngine, DecodeState

class V:
    def __len__(self): return 729
cfg = NS(MODEL=NS(num_layers=2, num_heads=8, units=512, inner_size=2048, dropout=0.0, attention_dropout=0.0, same_length=True, clamp_len=-1),
         TRAIN=NS(tgt_length=1, mem_length=2048))
torch.manual_seed(0)
m = MemTransformerLM(cfg, V()).cuda().eval()
with torch.no_grad():
    for n, p in m.named_parameters():
        p.normal_(1.0 if n.endswith("layer_norm.weight") else 0.0, 0.02)
eng = DecodeEngine(m, batch=64, mem_len=2048, same_length=True, precision="bf16")
for kc, vc in zip(eng.kc, eng.vc):
    kc.normal_(0, 0.5); vc.normal_(0, 0.5)
st = DecodeState(2048, 2047)
tok = torch.randint(1, 700, (64,), device="cuda")
for _ in range(4):
    lg, st = eng.step(tok, st)
    tok, _ = eng.sample(lg, 0.95, 0, 0.9, None, 1, 0)
torch.cuda.synchronize(); print("done")
