"""ORACLE -- test infrastructure, NOT product code.

CPU restatement (torch tensors used as the array library, fp32 or fp64) of the ComMU
Transformer-XL hot path: the algorithm of `commu/model/model.py` (MemTransformerLM) and of the
train-step tail of `train.py`, restated functionally from SURVEY.md section 3.3 / 8(a).  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may import
this module; the product path never does (it fails loudly without the CUDA library).

Parity pinning: the reference ships no golden vectors or tests, so this oracle is
pinned against outputs of the reference itself, imported in the build container and frozen under
tests/golden/*.npz by tests/golden/make_golden.py (tests/test_oracle_golden.py checks them).

Differences in formulation (same arithmetic result):
  * the relative shift (model.py:251-265) is an index:  BD[i, j] = (q_i + r_r_bias) . R[i + M - j]
    with R indexed by distance (model.py:578-583 builds pos_emb row jr for distance K-1-jr);
  * the attention mask (model.py:549-574) is an analytic predicate, never materialised per batch;
  * parameters are a flat dict keyed by the reference state_dict names.
"""
import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F


def make_cfg(n_layer, n_head, d_model, d_inner, tgt_len, mem_len, same_length=False, clamp_len=-1,
             n_token=729):
    return SimpleNamespace(n_layer=n_layer, n_head=n_head, d_model=d_model, d_inner=d_inner,
                           tgt_len=tgt_len, mem_len=mem_len, same_length=same_length,
                           clamp_len=clamp_len, n_token=n_token, d_head=d_model // n_head)


def param_names(cfg):
    names = ["r_w_bias", "r_r_bias", "word_emb.emb_layers.0.weight"]
    for i in range(cfg.n_layer):
        p = "layers.%d." % i
        names += [p + "dec_attn.qkv_net.weight", p + "dec_attn.o_net.weight",
                  p + "dec_attn.layer_norm.weight", p + "dec_attn.layer_norm.bias",
                  p + "dec_attn.r_net.weight",
                  p + "pos_ff.CoreNet.0.weight", p + "pos_ff.CoreNet.0.bias",
                  p + "pos_ff.CoreNet.3.weight", p + "pos_ff.CoreNet.3.bias",
                  p + "pos_ff.layer_norm.weight", p + "pos_ff.layer_norm.bias"]
    names += ["crit.out_layers.0.bias"]
    return names


def param_shapes(cfg):
    d, H, Dh, Di, V = cfg.d_model, cfg.n_head, cfg.d_head, cfg.d_inner, cfg.n_token
    shp = {"r_w_bias": (H, Dh), "r_r_bias": (H, Dh), "word_emb.emb_layers.0.weight": (V, d),
           "crit.out_layers.0.bias": (V,)}
    for i in range(cfg.n_layer):
        p = "layers.%d." % i
        shp[p + "dec_attn.qkv_net.weight"] = (3 * H * Dh, d)
        shp[p + "dec_attn.o_net.weight"] = (d, H * Dh)
        shp[p + "dec_attn.r_net.weight"] = (H * Dh, d)
        shp[p + "dec_attn.layer_norm.weight"] = (d,)
        shp[p + "dec_attn.layer_norm.bias"] = (d,)
        shp[p + "pos_ff.CoreNet.0.weight"] = (Di, d)
        shp[p + "pos_ff.CoreNet.0.bias"] = (Di,)
        shp[p + "pos_ff.CoreNet.3.weight"] = (d, Di)
        shp[p + "pos_ff.CoreNet.3.bias"] = (d,)
        shp[p + "pos_ff.layer_norm.weight"] = (d,)
        shp[p + "pos_ff.layer_norm.bias"] = (d,)
    return shp


def init_params(cfg, seed, std=0.01, dtype=torch.float32):
    """Recipe of train.py:291-342: N(0, std) for every matrix / embedding / r_*_bias, zero biases,
    LayerNorm gain ~ N(1, std).  (Different RNG consumption order than the reference, so use the
    golden files when the exact reference tensors are needed.)"""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_shapes(cfg).items():
        if name.endswith("layer_norm.weight"):
            t = 1.0 + std * torch.randn(shape, generator=g)
        elif name.endswith(".bias") and "r_w_bias" not in name and "r_r_bias" not in name:
            t = torch.zeros(shape)
        else:
            t = std * torch.randn(shape, generator=g)
        out[name] = t.to(dtype)
    return out


def inv_freq(d_model, dtype=torch.float32):
    # model.py:142
    return 1.0 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model)).to(dtype)


def sinusoid_by_distance(klen, d_model, clamp_len, dtype, device=None):
    """Row delta holds [sin(delta*f) | cos(delta*f)] (model.py:145-147), delta clamped like
    model.py:581-582."""
    dist = torch.arange(klen, dtype=dtype, device=device)
    if clamp_len > 0:
        dist = dist.clamp(max=clamp_len)
    ang = torch.outer(dist, inv_freq(d_model, dtype).to(device))
    return torch.cat([ang.sin(), ang.cos()], dim=-1)


def visible_mask(T, M, mem_len, same_length, reset, device=None):
    """valid[b, i, j] -- True where query i may look at key j."""
    K = T + M
    i = torch.arange(T, device=device)[:, None]
    j = torch.arange(K, device=device)[None, :]
    ok = j <= i + M
    if same_length:
        mask_len = K - mem_len
        shift = T - mask_len if mask_len > 0 else T
        ok = ok & (j > i - shift)
    B = reset.shape[0]
    ok = ok[None].expand(B, T, K).clone()
    if reset.any():
        ok[reset.bool()] &= (j >= M)
    return ok


def _layer(cfg, P, pre, x, mem, pos, valid, drop=None, l=0):
    dr = drop if drop is not None else (lambda site, layer, t: t)
    T, B, d = x.shape
    H, Dh = cfg.n_head, cfg.d_head
    M = mem.shape[0]
    K = T + M
    cat = torch.cat([mem, x], 0) if M > 0 else x
    Wqkv = P[pre + "dec_attn.qkv_net.weight"]
    Wq, Wk, Wv = Wqkv[: H * Dh], Wqkv[H * Dh: 2 * H * Dh], Wqkv[2 * H * Dh:]
    q = (x @ Wq.t()).view(T, B, H, Dh)
    k = (cat @ Wk.t()).view(K, B, H, Dh)
    v = (cat @ Wv.t()).view(K, B, H, Dh)
    R = (pos @ P[pre + "dec_attn.r_net.weight"].t()).view(K, H, Dh)      # by distance
    u, vb = P["r_w_bias"], P["r_r_bias"]
    AC = torch.einsum("ibhd,jbhd->bhij", q + u, k)
    QR = torch.einsum("ibhd,thd->bhit", q + vb, R)                        # [B,H,T,K] by distance
    ii = torch.arange(T, device=x.device)[:, None]
    jj = torch.arange(K, device=x.device)[None, :]
    dist = (ii + M - jj).clamp(min=0)                                      # masked where < 0
    BD = QR.gather(3, dist[None, None].expand(B, H, T, K))
    score = (AC + BD) * (1.0 / math.sqrt(Dh))
    score = score.masked_fill(~valid[:, None], float("-inf"))
    prob = dr("att", l, torch.softmax(score, dim=-1))                     # self.dropatt (model.py:337)
    av = torch.einsum("bhij,jbhd->ibhd", prob, v).reshape(T, B, H * Dh)
    a_out = dr("attn_out", l, av @ P[pre + "dec_attn.o_net.weight"].t())   # self.drop (model.py:349)
    y = F.layer_norm(x + a_out, (d,), P[pre + "dec_attn.layer_norm.weight"],
                     P[pre + "dec_attn.layer_norm.bias"], 1e-5)
    h = dr("ff_hid", l, torch.relu(y @ P[pre + "pos_ff.CoreNet.0.weight"].t() + P[pre + "pos_ff.CoreNet.0.bias"]))
    f = dr("ff_out", l, h @ P[pre + "pos_ff.CoreNet.3.weight"].t() + P[pre + "pos_ff.CoreNet.3.bias"])   # model.py:163-169
    return F.layer_norm(y + f, (d,), P[pre + "pos_ff.layer_norm.weight"],
                        P[pre + "pos_ff.layer_norm.bias"], 1e-5)


def hidden_forward(cfg, P, data, reset=None, mems=None, drop=None):
    """Returns (hidden [T,B,d], new_mems [L+1, min(M+T, mem_len), B, d] or None).
    drop: optional callable (site, layer, tensor) -> tensor standing in for the reference's nn.Dropout modules
    in training mode (sites "emb", "pos", "att", "attn_out", "ff_hid", "ff_out", "final"); the caller owns the
    masks, which lets the tests hand the oracle the very masks the CUDA kernels use."""
    T, B = data.shape
    dtype = P["r_w_bias"].dtype
    E = P["word_emb.emb_layers.0.weight"]
    x = E[data] * math.sqrt(cfg.d_model)
    M = 0 if (mems is None or mems.numel() == 0) else mems.shape[1]
    dev = E.device                      # the arrays live wherever the parameters do (CPU for parity tests and the
                                        # CPU baseline; bench.py's reference-GPU leg puts them on the B200)
    if reset is None:
        reset = torch.zeros(B, dtype=torch.bool, device=dev)
    valid = visible_mask(T, M, cfg.mem_len, cfg.same_length, reset, device=dev)
    pos = sinusoid_by_distance(T + M, cfg.d_model, cfg.clamp_len, dtype, device=dev)
    if drop is not None:                                                   # model.py:585-586
        x = drop("emb", 0, x)
        pos = drop("pos", 0, pos)
    hids = [x]
    for l in range(cfg.n_layer):
        mem = mems[l] if M > 0 else x.new_zeros(0, B, cfg.d_model)
        x = _layer(cfg, P, "layers.%d." % l, x, mem, pos, valid, drop, l)
        hids.append(x)
    new_mems = None
    if cfg.mem_len > 0:
        with torch.no_grad():
            stacked = torch.stack([h.detach() for h in hids])
            allh = torch.cat([mems, stacked], 1) if M > 0 else stacked
            end = M + T
            beg = max(0, end - cfg.mem_len)
            new_mems = allh[:, beg:end].clone()
    if drop is not None:
        x = drop("final", 0, x)                                            # model.py:600 (memory keeps the undropped rows)
    return x, new_mems


def logits_of(cfg, P, hidden):
    return hidden @ P["word_emb.emb_layers.0.weight"].t() + P["crit.out_layers.0.bias"]


def forward_loss(cfg, P, data, target, reset=None, mems=None, drop=None):
    """MemTransformerLM.forward (model.py:678-693): per-token NLL [T,B] (pad not masked)."""
    hidden, new_mems = hidden_forward(cfg, P, data, reset, mems, drop)
    T, B = target.shape
    lg = logits_of(cfg, P, hidden[-T:]).reshape(T * B, -1)
    nll = -torch.log_softmax(lg, dim=-1).gather(1, target.reshape(-1, 1)).squeeze(1)
    return nll.view(T, B), new_mems


def forward_generate(cfg, P, data, mems=None):
    """MemTransformerLM.forward_generate (model.py:606-628): raw logits [T,B,V]."""
    hidden, new_mems = hidden_forward(cfg, P, data, None, mems)
    return logits_of(cfg, P, hidden), new_mems


# ------------------------------------------------------------------------------------------------
# train-step tail (train.py:133-169, 441-461)
# ------------------------------------------------------------------------------------------------
def lr_multiplier(step, warmup_step, lr, lr_min):
    if step == 0 and warmup_step == 0:
        return 1.0
    if step > warmup_step:
        return max((warmup_step ** 0.5) / (step ** 0.5), lr_min / lr)
    return step / warmup_step


class AdamState:
    def __init__(self, P):
        self.m = {k: torch.zeros_like(v) for k, v in P.items()}
        self.v = {k: torch.zeros_like(v) for k, v in P.items()}
        self.t = 0


def train_step(cfg, P, opt, batches, mems, lr, clip=1.0, betas=(0.9, 0.999), eps=1e-8,
               pad_id=0, world=1, drop=None):
    """One optimizer step over `batches` = list of (data, target, reset) micro-batches
    (the reference's batch_chunk loop).  Updates P / opt in place.
    Returns (sum of the per-chunk mean losses (already / n_chunks), grad_norm, new mems list)."""
    names = list(P.keys())
    leaves = {k: P[k].detach().clone().requires_grad_(True) for k in names}
    total = 0.0
    n_chunks = len(batches)
    new_mems = []
    for c, (data, target, reset) in enumerate(batches):
        nll, nm = forward_loss(cfg, leaves, data, target, reset, mems[c], drop=drop)
        loss = nll[target != pad_id].float().mean() / n_chunks
        loss.backward()
        total += float(loss.detach())
        new_mems.append(nm)
    grads = {k: (leaves[k].grad if leaves[k].grad is not None else torch.zeros_like(P[k])) / world
             for k in names}
    gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).float()
    coef = min(1.0, float(clip / (gnorm + 1e-6)))
    opt.t += 1
    bc1 = 1.0 - betas[0] ** opt.t
    bc2 = 1.0 - betas[1] ** opt.t
    for k in names:
        g = grads[k] * coef
        opt.m[k].mul_(betas[0]).add_(g, alpha=1 - betas[0])
        opt.v[k].mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        denom = opt.v[k].sqrt() / math.sqrt(bc2) + eps
        P[k] = P[k] - (lr / bc1) * opt.m[k] / denom
    return total, float(gnorm), new_mems, grads


# ------------------------------------------------------------------------------------------------
# sampler (commu/midi_generator/midi_inferrer.py:199-237), stated for one row of logits
# ------------------------------------------------------------------------------------------------
def sampler_probs(logits_full, temperature, top_k=0, top_p=0.0, wrong_tokens=()):
    """logits_full: [V] raw logits (token 0 included).  Returns probs [V] after the reference's
    calc_probs + apply_sampling: token 0 is never sampled (:206, :220); temperature 0 = one-hot
    argmax (:211-213); top-k keeps the k largest probabilities (:224-226); `wrong_tokens` are
    zeroed (:227-229); renormalised (:230-231).  top_p (not in the reference; BASELINE config 4):
    keep the smallest prefix of the descending-sorted distribution whose mass reaches top_p."""
    lg = logits_full[1:].double() if logits_full.dtype == torch.float64 else logits_full[1:].float()
    if temperature == 0:
        probs = torch.zeros_like(lg)
        probs[lg.argmax()] = 1.0
    else:
        probs = torch.softmax(lg / temperature, dim=-1)
    probs = F.pad(probs, [1, 0])
    mask = torch.zeros_like(probs)
    if top_k and top_k > 0:
        _, idx = torch.topk(probs, top_k)
        mask[idx] = 1.0
    else:
        mask[:] = 1.0
    if top_p and top_p > 0.0:
        sp, si = torch.sort(probs, descending=True, stable=True)
        csum = torch.cumsum(sp, 0)
        keep = (csum - sp) < top_p          # keep tokens whose preceding mass is < top_p
        pm = torch.zeros_like(probs)
        pm[si[keep]] = 1.0
        mask = mask * pm
    for w in wrong_tokens:
        mask[w] = 0.0
    probs = probs * mask
    return probs / probs.sum()


def greedy_token(logits_full):
    """Q4 of SURVEY.md 3.2: 1 + argmax(logits[1:])."""
    return 1 + int(torch.argmax(logits_full[1:]))
