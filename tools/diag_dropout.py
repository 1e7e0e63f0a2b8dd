"""Diagnostic: per-parameter gradient error of the native model vs the oracle, with and without dropout."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "commu-code_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import orc, attn_keep_mask, drop_keep_mask, drop_keep_prob
from test_dropout_gpu import _tiny_model
from commu.engine import native_lm as nl

for (p_d, p_a) in ((0.0, 0.0), (0.1, 0.0), (0.0, 0.1), (0.1, 0.1)):
    m = _tiny_model(p_d, p_a)
    L, H, d, Di, T, B, V = 2, 2, 128, 256, 64, 3, 97
    cfg = orc.make_cfg(n_layer=L, n_head=H, d_model=d, d_inner=Di, tgt_len=T, mem_len=64, n_token=V)
    P = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.state_dict().items()
         if k not in ("crit.out_layers.0.weight", "pos_emb.inv_freq")}
    g = torch.Generator().manual_seed(2)
    m.train()
    mems_n = mems_o = None
    for seg in range(2):
        data = torch.randint(1, V, (T, B), generator=g); target = torch.randint(1, V, (T, B), generator=g)
        torch.manual_seed(100 + seg); base = int(torch.randint(0, 2 ** 62, (1,)).item())
        torch.manual_seed(100 + seg)
        ln, mems_n = m(data.cuda(), target.cuda(), None, mems_n)
        M = 0 if seg == 0 else 64; K = T + M
        sites = {"emb": nl.SITE_EMB, "pos": nl.SITE_POS, "att": nl.SITE_ATT, "attn_out": nl.SITE_ATTN_OUT,
                 "ff_hid": nl.SITE_FF_HID, "ff_out": nl.SITE_FF_OUT, "final": nl.SITE_FINAL}
        def drop(site, layer, t):
            seed = nl.site_seed(base, layer, sites[site])
            if site == "att":
                if p_a <= 0: return t
                return t * attn_keep_mask(seed, B, H, T, K, p_a) / drop_keep_prob(p_a)
            if p_d <= 0: return t
            if site == "pos": keep = drop_keep_mask(seed, K, d, p_d)
            else:
                c = t.shape[-1]; keep = drop_keep_mask(seed, T * B, c, p_d).view(T, B, c)
            return t * keep / drop_keep_prob(p_d)
        lo, mems_o = orc.forward_loss(cfg, P, data, target, None, mems_o, drop=drop)
        print("p", p_d, p_a, "seg", seg, "loss rel", abs(float(ln.mean()) - float(lo.mean())) / float(lo.mean()))
        ln.mean().backward(); lo.mean().backward()
    errs = []
    for name, prm in m.named_parameters():
        if name == "crit.out_layers.0.weight" or prm.grad is None: continue
        go = P[name].grad; gn = prm.grad.cpu()
        errs.append((float((gn - go).abs().max() / (go.abs().max() + 1e-12)), name))
    errs.sort(reverse=True)
    print("  worst:", [(round(e, 4), n) for e, n in errs[:6]])
