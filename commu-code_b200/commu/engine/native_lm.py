"""Host-side orchestration of the native (sm_100a) Transformer-XL forward / backward.

Python here is tensor plumbing only: buffer allocation (PyTorch caching allocator), the bf16 weight
shadow bookkeeping and the order of C-ABI calls.  All arithmetic runs in libcommu_b200.so.

Replaces the body of MemTransformerLM._forward / forward / forward_generate
(reference commu/model/model.py:540-693) and their autograd backward.

Layout conventions
  * activations are [T*B, ld] row-major with row = i*B + b (the reference's [T,B,*] flattened);
  * d_model, d_inner and the vocabulary are padded to multiples of 64 in BUFFERS (dp, dip, vp), the
    head dim is padded to 64 (hd = H*64); padding columns are kept at zero;
  * memory ("mems") lives as bf16 [M*B, dp] per layer -- exactly the operand the K/V projection
    consumes -- wrapped in a `Mems` handle that mimics the reference tensor's surface.
"""
import math
import os

import torch

from commu import _native as nv


def _ceil(a, b):
    return (a + b - 1) // b * b


_M64 = (1 << 64) - 1


def _mix64(x):
    """splitmix64 finaliser: seeds of the individual dropout sites from (base seed, layer, site)."""
    x = (x + 0x9E3779B97F4A7C15) & _M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _M64
    return x ^ (x >> 31)


# dropout sites (reference model.py): embedding :585, positional table :586, attention probabilities :337,
# attention output :349, FF hidden :166, FF output :168, final hidden :600
SITE_EMB, SITE_POS, SITE_ATT, SITE_ATTN_OUT, SITE_FF_HID, SITE_FF_OUT, SITE_FINAL = range(7)


def site_seed(base, layer, site):
    return _mix64((base * 0x2545F4914F6CDD1D + layer * 16 + site) & _M64)


def keep_prob(p):
    """Effective keep probability of the kernels' 15-bit threshold (dropout.cuh)."""
    thr = min(32767, max(0, int(p * 32768.0 + 0.5)))
    return 1.0 - thr / 32768.0


class Mems:
    """Opaque stand-in for the reference's `mems` tensor [L+1, M, B, d] (model.py:498-538).
    Callers only store it and pass it back; `.shape`, `.size()`, `len()`, `[i]` and `.float()` work."""

    def __init__(self, bufs, mlen, B, d, dp):
        self.bufs = bufs        # list of L+1 bf16 tensors [mlen*B, dp]
        self.mlen, self.B, self.d, self.dp = mlen, B, d, dp

    @property
    def shape(self):
        return torch.Size((len(self.bufs), self.mlen, self.B, self.d))

    def size(self, k=None):
        return self.shape if k is None else self.shape[k]

    def dim(self):
        return 4

    def numel(self):
        return len(self.bufs) * self.mlen * self.B * self.d

    def __len__(self):
        return len(self.bufs)

    def __getitem__(self, i):
        return self.bufs[i].view(self.mlen, self.B, self.dp)[:, :, :self.d].float()

    def detach(self):
        return self

    def float(self):
        return torch.stack([self[i] for i in range(len(self.bufs))])

    to_tensor = float


class NativeLM:
    """Stateless-ish engine bound to a set of fp32 master parameters (reference state_dict names)."""

    def __init__(self, params, n_layer, n_head, d_model, d_inner, n_token, inv_freq):
        self.P = params
        self.L, self.H, self.d, self.Di, self.V = n_layer, n_head, d_model, d_inner, n_token
        self.Dh = d_model // n_head
        if self.Dh > 64:
            raise RuntimeError("commu_b200: head dim %d > 64 is not supported by the attention kernels" % self.Dh)
        self.dp, self.dip, self.vp = _ceil(d_model, 64), _ceil(d_inner, 64), _ceil(n_token, 64)
        self.hd = n_head * 64
        self.inv_freq = inv_freq
        self.dev = params["r_w_bias"].device
        if self.dev.type != "cuda":
            raise RuntimeError("commu_b200: the native engine needs CUDA tensors (no CPU fallback)")
        nv.lib()  # fail loudly if the shared library is missing
        self.aligned = (self.Dh == 64 and self.d == self.dp and self.Di == self.dip)
        self._shadow = None
        self._shadow_version = None
        self._pos_cache = {}
        # attention forward implementation: "tc" = tcgen05 / TMEM kernel, "v1" = warp-MMA kernel
        self.attn_fwd_impl = {"tc": "commu_relattn_fwd_tc", "v1": "commu_relattn_fwd"}[
            os.environ.get("COMMU_ATTN_FWD", "tc")]
        # attention backward: "mat" = probabilities stored by the forward, dS materialised once, band GEMMs
        # (attn_bwd_mat.cu); "recompute" = the three recompute passes (no extra memory)
        self.bwd_materialise = (os.environ.get("COMMU_ATTN_BWD", "mat") == "mat"
                                and self.attn_fwd_impl == "commu_relattn_fwd_tc")
        self.saved = None

    # ------------------------------------------------------------------ weight shadows ------------
    def _alloc_shadow(self):
        dev, bf = self.dev, torch.bfloat16
        S = {"layers": []}
        for _ in range(self.L):
            S["layers"].append(dict(
                wq=torch.zeros(self.hd, self.dp, device=dev, dtype=bf),
                wkv=torch.zeros(2 * self.hd, self.dp, device=dev, dtype=bf),
                wr=torch.zeros(self.hd, self.dp, device=dev, dtype=bf),
                wo=torch.zeros(self.dp, self.hd, device=dev, dtype=bf),
                w1=torch.zeros(self.dip, self.dp, device=dev, dtype=bf),
                w2=torch.zeros(self.dp, self.dip, device=dev, dtype=bf),
                b1=torch.zeros(self.dip, device=dev), b2=torch.zeros(self.dp, device=dev)))
        S["emb"] = torch.zeros(self.V, self.dp, device=dev, dtype=bf)
        S["u"] = torch.zeros(self.H, 64, device=dev)
        S["vb"] = torch.zeros(self.H, 64, device=dev)
        S["lbias"] = torch.zeros(self.vp, device=dev)
        return S

    @torch.no_grad()
    def adopt_shadow(self, flat_bf16, flat_f32, offsets):
        """Aligned shapes only (Dh = 64, d and d_inner multiples of 64): the operand shadows become VIEWS of a flat bf16
        arena that mirrors the trainer's flat fp32 master arena (same offsets), so the Adam kernel can emit them
        (commu_clip_adam p_bf16) and the 85 cast launches after every optimizer step disappear."""
        assert self.aligned
        H, d, Di, V = self.H, self.d, self.Di, self.V
        view = lambda name, rows, cols, skip=0: flat_bf16[offsets[name] + skip: offsets[name] + skip + rows * cols].view(rows, cols)
        S = {"layers": []}
        for l in range(self.L):
            pre = "layers.%d." % l
            S["layers"].append(dict(
                wq=view(pre + "dec_attn.qkv_net.weight", self.hd, d),
                wkv=view(pre + "dec_attn.qkv_net.weight", 2 * self.hd, d, skip=self.hd * d),
                wr=view(pre + "dec_attn.r_net.weight", self.hd, d),
                wo=view(pre + "dec_attn.o_net.weight", d, self.hd),
                w1=view(pre + "pos_ff.CoreNet.0.weight", Di, d),
                w2=view(pre + "pos_ff.CoreNet.3.weight", d, Di),
                b1=self.P[pre + "pos_ff.CoreNet.0.bias"], b2=self.P[pre + "pos_ff.CoreNet.3.bias"]))   # fp32 masters themselves
        S["emb"] = view("word_emb.emb_layers.0.weight", V, d)
        S["u"], S["vb"] = self.P["r_w_bias"], self.P["r_r_bias"]
        S["lbias"] = torch.zeros(self.vp, device=self.dev)
        self._shadow = S
        self._ext = (flat_bf16, flat_f32)
        self._shadow_version = None

    @torch.no_grad()
    def refresh_shadow(self, after_update=False):
        """fp32 masters -> bf16 operand shadows (padded).  Call after every optimizer step."""
        if getattr(self, "_ext", None) is not None:
            # adopted arena: after an optimizer step the Adam kernel has already written the bf16 copy; any other
            # change of the masters (load_state_dict, manual edits) is caught by the version check -> one flat cast
            S, P = self._shadow, self.P
            if S["layers"][0]["b1"] is not P["layers.0.pos_ff.CoreNet.0.bias"]:   # parameters re-materialised
                self._ext = None
                self._shadow = None
                return self.refresh_shadow()
            if not after_update:
                self._ext[0].copy_(self._ext[1])
            S["lbias"][:self.V].copy_(P["crit.out_layers.0.bias"])
            self._shadow_version = self._param_version()
            return
        if self._shadow is None:
            self._shadow = self._alloc_shadow()
        S, P = self._shadow, self.P
        H, Dh, d, Di = self.H, self.Dh, self.d, self.Di
        for l in range(self.L):
            pre = "layers.%d." % l
            s = S["layers"][l]
            wqkv = P[pre + "dec_attn.qkv_net.weight"]
            nv.call("commu_cast_pad", wqkv, d, H * Dh, d, Dh, 64, d, self.dp, s["wq"], self.dp, 0)
            nv.call("commu_cast_pad", wqkv[H * Dh:], d, 2 * H * Dh, d, Dh, 64, d, self.dp, s["wkv"], self.dp, 0)
            nv.call("commu_cast_pad", P[pre + "dec_attn.r_net.weight"], d, H * Dh, d, Dh, 64, d, self.dp,
                    s["wr"], self.dp, 0)
            nv.call("commu_cast_pad", P[pre + "dec_attn.o_net.weight"], H * Dh, d, H * Dh, d, self.dp, Dh, 64,
                    s["wo"], self.hd, 0)
            nv.call("commu_cast_pad", P[pre + "pos_ff.CoreNet.0.weight"], d, Di, d, Di, self.dip, d, self.dp,
                    s["w1"], self.dp, 0)
            nv.call("commu_cast_pad", P[pre + "pos_ff.CoreNet.3.weight"], Di, d, Di, d, self.dp, Di, self.dip,
                    s["w2"], self.dip, 0)
            s["b1"][:Di].copy_(P[pre + "pos_ff.CoreNet.0.bias"])
            s["b2"][:d].copy_(P[pre + "pos_ff.CoreNet.3.bias"])
        nv.call("commu_cast_pad", P["word_emb.emb_layers.0.weight"], d, self.V, d, self.V, self.V, d, self.dp,
                S["emb"], self.dp, 0)
        S["u"][:, :Dh].copy_(P["r_w_bias"])
        S["vb"][:, :Dh].copy_(P["r_r_bias"])
        S["lbias"][:self.V].copy_(P["crit.out_layers.0.bias"])
        self._shadow_version = self._param_version()

    def _param_version(self):
        return tuple(p._version for p in self.P.values())

    def shadow(self):
        if self._shadow is None or self._shadow_version != self._param_version():
            self.refresh_shadow()
        return self._shadow

    def pos_table(self, klen, clamp_len):
        key = (klen, clamp_len)
        t = self._pos_cache.get(key)
        if t is None:
            t = torch.empty(klen, self.dp, device=self.dev, dtype=torch.bfloat16)
            nv.call("commu_pos_table", self.inv_freq, klen, clamp_len, self.d, self.dp, t, None)
            if len(self._pos_cache) > 16:
                self._pos_cache.clear()
            self._pos_cache[key] = t
        return t

    # ------------------------------------------------------------------ forward --------------------
    @torch.no_grad()
    def _drop(self, x, rows, cols, p, seed, res=None, out_f32=None, out_bf16=None):
        """out = res + keep(x) / (1 - p)   (commu_dropout; x fp32 or bf16, in place allowed)."""
        nv.call("commu_dropout", x, int(x.dtype == torch.bfloat16), x.stride(0), res,
                res.stride(0) if res is not None else 0, rows, cols, float(p), seed,
                out_f32, out_f32.stride(0) if out_f32 is not None else 0,
                out_bf16, out_bf16.stride(0) if out_bf16 is not None else 0)

    def hidden_forward(self, data, reset, mems, mem_len, same_length, clamp_len, save, dropout=None):
        """Runs embedding + L layers.  Returns (xL_f32 [T*B, dp], xL_bf16, new_mems, ctx).
        dropout: None or (p, p_att, base_seed) - the reference's nn.Dropout sites in training mode."""
        T, B = data.shape
        rows = T * B
        dev, bf = self.dev, torch.bfloat16
        S = self.shadow()
        M = 0 if mems is None else mems.mlen
        if M > 0 and (mems.B != B or len(mems.bufs) != self.L + 1):
            raise RuntimeError("commu_b200: mems do not match this batch (B=%d vs %d)" % (mems.B, B))
        K = T + M
        krows = K * B
        mask_len = K - mem_len
        shift = (T - mask_len) if mask_len > 0 else T
        reset_u8 = None
        if reset is not None:
            reset_u8 = reset.to(device=dev, dtype=torch.uint8).contiguous()
        pos = self.pos_table(K, clamp_len)
        scale = 1.0 / math.sqrt(self.Dh)
        tok = data.reshape(-1).contiguous()
        pd, patt, dbase = dropout if dropout is not None else (0.0, 0.0, 0)
        if patt > 0 and self.attn_fwd_impl != "commu_relattn_fwd_tc":
            raise RuntimeError("commu_b200: attention dropout needs the tcgen05 attention kernels (COMMU_ATTN_FWD=tc)")
        if pd > 0:   # pos_emb = self.drop(pos_emb)  (model.py:586)
            pos_d = torch.empty_like(pos)
            self._drop(pos, K, self.dp, pd, site_seed(dbase, 0, SITE_POS), out_bf16=pos_d)
            pos = pos_d

        def new_cat(l):
            c = torch.empty(krows, self.dp, device=dev, dtype=bf)
            if M > 0:
                c[: M * B].copy_(mems.bufs[l])
            return c

        cats = [new_cat(0)]
        x = torch.empty(rows, self.dp, device=dev)
        nv.call("commu_embed_fwd", tok, self.P["word_emb.emb_layers.0.weight"], self.d, self.dp,
                math.sqrt(self.d), rows, x, self.dp, cats[0][M * B:], self.dp)
        if pd > 0:   # core_out = self.drop(word_emb)  (model.py:585); the dropped embedding is what enters the memory
            self._drop(x, rows, self.dp, pd, site_seed(dbase, 0, SITE_EMB), out_f32=x, out_bf16=cats[0][M * B:])
        layers_ctx = []
        if save and self.bwd_materialise:
            p_bytes, mt_bytes, _, _ = nv.attn_sizes(T, M, B, self.H)
        for l in range(self.L):
            s = S["layers"][l]
            cat = cats[l]
            xb = cat[M * B:]
            q = torch.empty(rows, self.hd, device=dev, dtype=bf)
            nv.gemm(xb, s["wq"], m=rows, n=self.hd, k=self.dp, out_bf16=q)
            kv = torch.empty(krows, 2 * self.hd, device=dev, dtype=bf)
            nv.gemm(cat, s["wkv"], m=krows, n=2 * self.hd, k=self.dp, out_bf16=kv)
            r = torch.empty(K, self.hd, device=dev, dtype=bf)
            nv.gemm(pos, s["wr"], m=K, n=self.hd, k=self.dp, out_bf16=r)
            av = torch.empty(rows, self.hd, device=dev, dtype=bf)
            lse = torch.empty(B, self.H, T, device=dev)
            qu = torch.empty(rows, self.hd, device=dev, dtype=bf) if save else None
            qv = torch.empty(rows, self.hd, device=dev, dtype=bf) if save else None
            nv.call("commu_relattn_set_dropout", float(patt), site_seed(dbase, l, SITE_ATT))
            psv = mtv = None
            if self.attn_fwd_impl == "commu_relattn_fwd_tc":
                if save and self.bwd_materialise:
                    # probabilities kept for the backward (2 bytes per score element; attn_bwd_mat.cu); never
                    # initialised: the backward only reads the tiles the forward wrote
                    psv = torch.empty(p_bytes, dtype=torch.uint8, device=dev)
                    mtv = torch.empty(mt_bytes // 4, device=dev)
                extra = (psv, mtv)
            else:
                extra = ()
            nv.call(self.attn_fwd_impl, q, self.hd, kv, kv[:, self.hd:], 2 * self.hd, r, self.hd, K,
                    S["u"], S["vb"], reset_u8, T, M, B, self.H, int(bool(same_length)), shift, scale,
                    av, self.hd, lse, qu, qv, *extra)
            z1 = torch.empty(rows, self.dp, device=dev)
            # attn_out = self.drop(self.o_net(attn_vec)); w + attn_out  (model.py:348-352): dropout and residual are the
            # GEMM epilogue
            nv.gemm(av, s["wo"], m=rows, n=self.dp, k=self.hd, add_f32=x, out_f32=z1, drop_p=pd,
                    drop_seed=site_seed(dbase, l, SITE_ATTN_OUT) if pd > 0 else 0)
            pre = "layers.%d." % l
            y1 = torch.empty(rows, self.dp, device=dev)
            y1b = torch.empty(rows, self.dp, device=dev, dtype=bf)
            mean1 = torch.empty(rows, device=dev)
            rstd1 = torch.empty(rows, device=dev)
            nv.call("commu_layernorm_fwd", z1, self.dp, self.P[pre + "dec_attn.layer_norm.weight"],
                    self.P[pre + "dec_attn.layer_norm.bias"], self.d, self.dp, 1e-5, rows, y1, self.dp,
                    y1b, self.dp, mean1, rstd1)
            hdn = torch.empty(rows, self.dip, device=dev, dtype=bf)
            # Linear, ReLU, Dropout, Linear, Dropout; inp + core_out  (model.py:163-179): both dropouts in the epilogues
            nv.gemm(y1b, s["w1"], m=rows, n=self.dip, k=self.dp, bias=s["b1"], relu=True, out_bf16=hdn, drop_p=pd,
                    drop_seed=site_seed(dbase, l, SITE_FF_HID) if pd > 0 else 0)
            z2 = torch.empty(rows, self.dp, device=dev)
            nv.gemm(hdn, s["w2"], m=rows, n=self.dp, k=self.dip, bias=s["b2"], add_f32=y1, out_f32=z2, drop_p=pd,
                    drop_seed=site_seed(dbase, l, SITE_FF_OUT) if pd > 0 else 0)
            nxt = new_cat(l + 1) if l + 1 < self.L else torch.empty(krows, self.dp, device=dev, dtype=bf)
            if l + 1 == self.L and M > 0:
                nxt[: M * B].copy_(mems.bufs[self.L])
            x_next = torch.empty(rows, self.dp, device=dev)
            mean2 = torch.empty(rows, device=dev)
            rstd2 = torch.empty(rows, device=dev)
            nv.call("commu_layernorm_fwd", z2, self.dp, self.P[pre + "pos_ff.layer_norm.weight"],
                    self.P[pre + "pos_ff.layer_norm.bias"], self.d, self.dp, 1e-5, rows, x_next, self.dp,
                    nxt[M * B:], self.dp, mean2, rstd2)
            cats.append(nxt)
            if save:
                layers_ctx.append(dict(kv=kv, r=r, av=av, lse=lse, qu=qu, qv=qv, psv=psv, mtv=mtv, z1=z1, mean1=mean1,
                                       rstd1=rstd1, y1b=y1b, hdn=hdn, z2=z2, mean2=mean2, rstd2=rstd2))
            x = x_next
        # new memory: the last mem_len positions of [old mem ; this segment] for every layer input
        new_mems = None
        if mem_len > 0:
            keep = min(K, mem_len)
            new_mems = Mems([c[(K - keep) * B:] for c in cats], keep, B, self.d, self.dp)
        nv.call("commu_relattn_set_dropout", 0.0, 0)
        ctx = None
        if save:
            ctx = dict(T=T, B=B, M=M, K=K, shift=shift, same_length=int(bool(same_length)), scale=scale,
                       reset_u8=reset_u8, pos=pos, tok=tok, cats=cats, layers=layers_ctx, drop=(pd, patt, dbase))
        xb_out = cats[self.L][M * B:]
        if pd > 0:   # core_out = self.drop(core_out) before the loss (model.py:600); the memory keeps the undropped rows
            xb_out = torch.empty(rows, self.dp, device=dev, dtype=bf)
            self._drop(x, rows, self.dp, pd, site_seed(dbase, 0, SITE_FINAL), out_bf16=xb_out)
        return x, xb_out, new_mems, ctx

    @torch.no_grad()
    def forward_loss(self, data, target, reset, mems, mem_len, same_length, clamp_len, save=True, dropout=None):
        T, B = data.shape
        rows = T * B
        xf, xb, new_mems, ctx = self.hidden_forward(data, reset, mems, mem_len, same_length, clamp_len, save,
                                                    dropout=dropout)
        S = self.shadow()
        logits = torch.empty(rows, self.vp, device=self.dev)
        nv.gemm(xb, S["emb"], m=rows, n=self.V, k=self.dp, bias=S["lbias"], out_f32=logits)
        tgt = target.reshape(-1).contiguous()
        nll = torch.empty(rows, device=self.dev)
        lse = torch.empty(rows, device=self.dev)
        nv.call("commu_nll_fwd", logits, self.vp, self.V, tgt, rows, nll, lse)
        if save:
            ctx.update(logits=logits, lse_v=lse, tgt=tgt, xLb=xb)
            self.saved = ctx
        return nll.view(T, B), new_mems

    @torch.no_grad()
    def forward_logits(self, data, mems, mem_len, same_length, clamp_len):
        T, B = data.shape
        rows = T * B
        xf, xb, new_mems, _ = self.hidden_forward(data, None, mems, mem_len, same_length, clamp_len, False)
        S = self.shadow()
        logits = torch.empty(rows, self.vp, device=self.dev)
        nv.gemm(xb, S["emb"], m=rows, n=self.V, k=self.dp, bias=S["lbias"], out_f32=logits)
        return logits[:, :self.V].reshape(T, B, self.V), new_mems

    # ------------------------------------------------------------------ backward -------------------
    def _split_for(self, m, n, k):
        tiles = ((m + 127) // 128) * ((n + 255) // 256 if n > 128 else 1)
        kb = (k + 63) // 64
        s = max(1, min(kb, (2 * 148 + tiles - 1) // tiles))
        return s

    def _wgrad(self, a, b, m, n, k, grad, rseg=None, rseg_pad=None, cseg=None, cseg_pad=None):
        """grad[m_real, n_real] += a^T b with a [k, m_pad], b [k, n_pad] (both bf16, k = tokens)."""
        m_real, n_real = grad.shape
        direct = (m == m_real and n >= n_real and (rseg is None or rseg == rseg_pad)
                  and (cseg is None or cseg == cseg_pad))
        if direct:
            nv.gemm(a, b, m=m, n=n_real, k=k, a_mn=True, b_mn=True, split_k=self._split_for(m, n_real, k),
                    out_f32=grad, ld_out_f32=grad.stride(0), f32_atomic=True)
            return
        tmp = torch.zeros(m, n, device=self.dev)
        nv.gemm(a, b, m=m, n=n, k=k, a_mn=True, b_mn=True, split_k=self._split_for(m, n, k), out_f32=tmp,
                f32_atomic=True)
        nv.call("commu_unpad_accum", tmp, n, m_real, n_real, rseg or m_real, rseg_pad or m_real,
                cseg or n_real, cseg_pad or n_real, grad, grad.stride(0), 1.0)

    @torch.no_grad()
    def backward(self, dloss, grads, layer_done=None):
        """dloss: fp32 [T,B] gradient of the per-token NLL.  Accumulates (+=) into `grads`
        (dict: reference parameter name -> fp32 tensor of the parameter's shape).
        layer_done(l), if given, is called right after the last gradient of decoder layer l has been issued (layers
        run L-1 .. 0): the data-parallel trainer starts that layer's share of the gradient exchange there."""
        c = self.saved
        if c is None:
            raise RuntimeError("commu_b200: backward() without a saved forward")
        self.saved = None
        S = self.shadow()
        T, B, M, K = c["T"], c["B"], c["M"], c["K"]
        rows, krows = T * B, K * B
        dev, bf = self.dev, torch.bfloat16
        H, Dh, d, Di, V = self.H, self.Dh, self.d, self.Di, self.V
        dl = dloss.reshape(-1).contiguous().float()
        dlogits = torch.empty(rows, self.vp, device=dev, dtype=bf)
        nv.call("commu_nll_bwd", c["logits"], self.vp, V, self.vp, c["lse_v"], c["tgt"], dl, rows, dlogits,
                self.vp)
        g_emb = grads["word_emb.emb_layers.0.weight"]
        self._wgrad(dlogits, c["xLb"], V, self.dp, rows, g_emb)
        lb = torch.zeros(self.vp, device=dev)
        nv.call("commu_colsum_bf16", dlogits, self.vp, V, rows, lb)
        grads["crit.out_layers.0.bias"].add_(lb[:V])
        dx = torch.empty(rows, self.dp, device=dev)
        nv.gemm(dlogits, S["emb"], m=rows, n=self.dp, k=V, b_mn=True, out_f32=dx)
        pd, patt, dbase = c.get("drop", (0.0, 0.0, 0))
        if patt > 0 and not all(os.environ.get(e, "tc") == "tc" for e in ("COMMU_ATTN_BWD_DQ", "COMMU_ATTN_BWD_DKV", "COMMU_ATTN_BWD_DR")):
            raise RuntimeError("commu_b200: attention dropout needs the tcgen05 backward passes")
        if pd > 0:
            self._drop(dx, rows, self.dp, pd, site_seed(dbase, 0, SITE_FINAL), out_f32=dx)
        du = torch.zeros(H, 64, device=dev)
        dvb = torch.zeros(H, 64, device=dev)
        delta = torch.empty(B, H, T, device=dev)
        for l in range(self.L - 1, -1, -1):
            s = S["layers"][l]
            a = c["layers"][l]
            pre = "layers.%d." % l
            cat = c["cats"][l]
            xb = cat[M * B:]
            # ---- position-wise FF block ----
            dz2 = torch.empty(rows, self.dp, device=dev)
            dz2b = torch.empty(rows, self.dp, device=dev, dtype=bf)
            nv.call("commu_layernorm_bwd", dx, self.dp, a["z2"], self.dp, a["mean2"], a["rstd2"],
                    self.P[pre + "pos_ff.layer_norm.weight"], d, self.dp, rows, dz2, self.dp, dz2b, self.dp,
                    grads[pre + "pos_ff.layer_norm.weight"], grads[pre + "pos_ff.layer_norm.bias"],
                    # the bf16 copy carries the gradient through the FF output dropout; the residual branch keeps dz2
                    float(pd), site_seed(dbase, l, SITE_FF_OUT) if pd > 0 else 0)
            nv.call("commu_colsum_bf16", dz2b, self.dp, d, rows, grads[pre + "pos_ff.CoreNet.3.bias"])
            self._wgrad(dz2b, a["hdn"], self.dp, self.dip, rows, grads[pre + "pos_ff.CoreNet.3.weight"])
            dpre = torch.empty(rows, self.dip, device=dev, dtype=bf)
            # hidden dropout: the saved (dropped) hidden is > 0 exactly where ReLU and the mask passed
            nv.gemm(dz2b, s["w2"], m=rows, n=self.dip, k=self.dp, b_mn=True, relu_mask=a["hdn"], out_bf16=dpre,
                    alpha=(1.0 / keep_prob(pd)) if pd > 0 else 1.0)
            nv.call("commu_colsum_bf16", dpre, self.dip, Di, rows, grads[pre + "pos_ff.CoreNet.0.bias"])
            self._wgrad(dpre, a["y1b"], self.dip, self.dp, rows, grads[pre + "pos_ff.CoreNet.0.weight"])
            dy1 = torch.empty(rows, self.dp, device=dev)
            nv.gemm(dpre, s["w1"], m=rows, n=self.dp, k=self.dip, b_mn=True, add_f32=dz2, out_f32=dy1)
            # ---- attention block ----
            dz1 = dz2  # reuse
            dz1b = dz2b
            nv.call("commu_layernorm_bwd", dy1, self.dp, a["z1"], self.dp, a["mean1"], a["rstd1"],
                    self.P[pre + "dec_attn.layer_norm.weight"], d, self.dp, rows, dz1, self.dp, dz1b, self.dp,
                    grads[pre + "dec_attn.layer_norm.weight"], grads[pre + "dec_attn.layer_norm.bias"],
                    # the bf16 copy carries the gradient through the attention output dropout
                    float(pd), site_seed(dbase, l, SITE_ATTN_OUT) if pd > 0 else 0)
            self._wgrad(dz1b, a["av"], self.dp, self.hd, rows, grads[pre + "dec_attn.o_net.weight"],
                        cseg=Dh, cseg_pad=64)
            dav = torch.empty(rows, self.hd, device=dev, dtype=bf)
            nv.gemm(dz1b, s["wo"], m=rows, n=self.hd, k=self.dp, b_mn=True, out_bf16=dav)
            dq = torch.empty(rows, self.hd, device=dev, dtype=bf)
            dkv = torch.empty(krows, 2 * self.hd, device=dev, dtype=bf)
            dr = torch.zeros(K, self.hd, device=dev)
            nv.call("commu_relattn_set_dropout", float(patt), site_seed(dbase, l, SITE_ATT))
            ws = nv.attn_bwd_workspace(T, M, B, H, dev) if a.get("psv") is not None else None
            nv.call("commu_relattn_bwd", a["qu"], a["qv"], self.hd, a["kv"], a["kv"][:, self.hd:], 2 * self.hd,
                    a["r"], self.hd, K, c["reset_u8"], T, M, B, H, c["same_length"], c["shift"], c["scale"],
                    a["av"], self.hd, a["lse"], dav, self.hd, delta, dq, self.hd, dkv, dkv[:, self.hd:],
                    2 * self.hd, dr, du, dvb, a.get("psv"), a.get("mtv"), ws, ws.numel() if ws is not None else 0)
            a["psv"] = a["mtv"] = None      # 2 GB per layer at the benchmark shape: release as soon as consumed
            g_qkv = grads[pre + "dec_attn.qkv_net.weight"]
            self._wgrad(dq, xb, self.hd, self.dp, rows, g_qkv[: H * Dh], rseg=Dh, rseg_pad=64)
            self._wgrad(dkv, cat, 2 * self.hd, self.dp, krows, g_qkv[H * Dh:], rseg=Dh, rseg_pad=64)
            drb = torch.empty(K, self.hd, device=dev, dtype=bf)
            nv.call("commu_cast_pad", dr, self.hd, K, self.hd, K, K, self.hd, self.hd, drb, self.hd, 0)
            self._wgrad(drb, c["pos"], self.hd, self.dp, K, grads[pre + "dec_attn.r_net.weight"], rseg=Dh,
                        rseg_pad=64)
            # gradient of the layer input: residual + through Wq (all rows) + through Wkv (segment rows)
            nv.gemm(dq, s["wq"], m=rows, n=self.dp, k=self.hd, b_mn=True, add_f32=dz1, out_f32=dx)
            nv.gemm(dkv[M * B:], s["wkv"], m=rows, n=self.dp, k=2 * self.hd, b_mn=True, add_f32=dx, out_f32=dx)
            if layer_done is not None:
                layer_done(l)
        nv.call("commu_relattn_set_dropout", 0.0, 0)
        if pd > 0:
            self._drop(dx, rows, self.dp, pd, site_seed(dbase, 0, SITE_EMB), out_f32=dx)
        nv.call("commu_embed_bwd", c["tok"], dx, self.dp, d, math.sqrt(d), rows, g_emb)
        grads["r_w_bias"].add_(du[:, :Dh])
        grads["r_r_bias"].add_(dvb[:, :Dh])
