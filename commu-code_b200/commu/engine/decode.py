"""Native incremental decode engine: per-layer projected K/V ring cache + cached R-by-distance table,
one kernel sequence per generated token for a whole batch of sequences, on-device sampling.

Replaces the per-token `forward_generate` + `calc_probs` / `apply_sampling` / `infer_token` loop of
the reference (commu/midi_generator/midi_inferrer.py:199-237 over commu/model/model.py:606-628).
`precision="fp32"` streams fp32 weights / cache (greedy tokens bit-faithful to the fp32 reference up
to summation order); `precision="bf16"` halves the streamed bytes (throughput configuration).
"""
import math

import torch

from commu import _native as nv


class DecodeState:
    """Immutable handle playing the role of the reference's `mems` in the decode loop: how many
    tokens are cached and where the newest sits in the ring.  Passing an older handle back rewinds."""
    __slots__ = ("count", "slot")

    def __init__(self, count=0, slot=-1):
        self.count, self.slot = count, slot


class DecodeEngine:
    def __init__(self, model, batch, mem_len, same_length=True, precision="fp32"):
        if batch > 64:
            raise RuntimeError("commu_b200 decode: batch %d > 64 per engine (shard sequences across engines / GPUs)" % batch)
        self.m = model
        self.B, self.mem_len, self.same_length = batch, mem_len, bool(same_length)
        self.bf16 = precision == "bf16"
        if precision not in ("fp32", "bf16"):
            raise ValueError(precision)
        self.L, self.H, self.d, self.Dh = model.n_layer, model.n_head, model.d_model, model.d_head
        self.Di, self.V = model.d_inner, model.n_token
        if self.d % 4 or self.Di % 4:
            raise RuntimeError("commu_b200 decode: d_model and d_inner must be multiples of 4")
        self.C = mem_len + 1
        self.dev = model.r_w_bias.device
        if self.dev.type != "cuda":
            raise RuntimeError("commu_b200: decode needs CUDA (no CPU fallback)")
        nv.lib()
        self.scale = 1.0 / math.sqrt(self.Dh)
        self._prepare()

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _prepare(self):
        dev, m = self.dev, self.m
        cdt = torch.bfloat16 if self.bf16 else torch.float32
        H, Dh, d, C, B = self.H, self.Dh, self.d, self.C, self.B
        sd = {n: p for n, p in m.named_parameters()}
        self.W = []
        wcast = (lambda t: t.detach().to(torch.bfloat16).contiguous()) if self.bf16 else (lambda t: t.detach().contiguous())
        # sinusoid table by distance, fp32 [C, d]
        pos = torch.empty(C, d, device=dev)
        nv.call("commu_pos_table", m.pos_emb.inv_freq, C, int(m.clamp_len), d, d, None, pos)
        self.u = torch.zeros(H, 64, device=dev)
        self.vb = torch.zeros(H, 64, device=dev)
        self.u[:, :Dh].copy_(sd["r_w_bias"])
        self.vb[:, :Dh].copy_(sd["r_r_bias"])
        self.kc, self.vc, self.rt = [], [], []
        for l in range(self.L):
            pre = "layers.%d." % l
            wo = sd[pre + "dec_attn.o_net.weight"].detach()
            wo_p = torch.zeros(d, H * 64, device=dev)
            wo_p.view(d, H, 64)[:, :, :Dh].copy_(wo.view(d, H, Dh))
            self.W.append(dict(
                qkv=wcast(sd[pre + "dec_attn.qkv_net.weight"]), o=wcast(wo_p),
                w1=wcast(sd[pre + "pos_ff.CoreNet.0.weight"]), b1=sd[pre + "pos_ff.CoreNet.0.bias"].detach(),
                w2=wcast(sd[pre + "pos_ff.CoreNet.3.weight"]), b2=sd[pre + "pos_ff.CoreNet.3.bias"].detach(),
                g1=sd[pre + "dec_attn.layer_norm.weight"].detach(), be1=sd[pre + "dec_attn.layer_norm.bias"].detach(),
                g2=sd[pre + "pos_ff.layer_norm.weight"].detach(), be2=sd[pre + "pos_ff.layer_norm.bias"].detach()))
            self.kc.append(torch.zeros(B, H, C, 64, device=dev, dtype=cdt))
            self.vc.append(torch.zeros(B, H, C, 64, device=dev, dtype=cdt))
            # R[a] = r_net(pos[a]) for every distance, projected once (weights are frozen at inference)
            wr = wcast(sd[pre + "dec_attn.r_net.weight"])
            rflat = torch.empty(C, H * Dh, device=dev)
            for r0 in range(0, C, 64):
                n = min(64, C - r0)
                self._linear(pos[r0:r0 + n], wr, None, False, None, rflat[r0:r0 + n], n, H * Dh, d)
            rt = torch.zeros(C, H, 64, device=dev, dtype=cdt)
            nv.call("commu_pad_heads", rflat, H * Dh, 0, C, H, Dh, rt, int(self.bf16), H * 64, 64, 0, None)
            self.rt.append(rt)
        self.emb32 = sd["word_emb.emb_layers.0.weight"].detach()
        self.emb = wcast(self.emb32)
        self.lbias = sd["crit.out_layers.0.bias"].detach()
        # step workspaces
        f = lambda *s: torch.empty(*s, device=dev)
        self.ws = dict(x=f(B, d), qkv=f(B, 3 * H * Dh), q=f(B, H, 64), att=f(B, H * 64), z=f(B, d), y=f(B, d),
                       h=f(B, self.Di), logits=f(B, self.V))
        # bf16 throughput mode: the linear layers run on the tcgen05 GEMM (weights streamed once per step,
        # the 64 batch rows are one half-filled 128-row MMA tile), so activations also exist as bf16 operands
        self.tc_linear = self.bf16 and d % 8 == 0 and self.Di % 8 == 0 and (H * Dh) % 8 == 0
        if self.tc_linear:
            hb = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
            self.wsb = dict(x=hb(B, d), att=hb(B, H * 64), y=hb(B, d), h=hb(B, self.Di))

    def _linear(self, x, w, bias, relu, res, out, B, N, K):
        nv.call("commu_decode_linear", x, x.stride(0), w, w.stride(0), int(w.dtype == torch.bfloat16), bias,
                int(relu), res, res.stride(0) if res is not None else 0, out, out.stride(0), B, N, K)

    def _ln(self, z, g, b, out):
        nv.call("commu_layernorm_fwd", z, self.d, g, b, self.d, self.d, 1e-5, self.B, out, self.d, None, 0,
                None, None)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, tokens, state):
        """tokens: int64 [B] on device.  Returns (logits fp32 [B, V] (a reused workspace), new state)."""
        slot = (state.slot + 1) % self.C
        cached = min(state.count, self.mem_len)
        n_vis = min(cached + 1, self.mem_len + (0 if self.same_length else 1))
        self._step_kernels(tokens, slot, n_vis, None)
        return self.ws["logits"], DecodeState(min(state.count + 1, self.mem_len), slot)

    def _step_kernels(self, tokens, slot, n_vis, dstate):
        """One token for every sequence.  With `dstate` (device int32[4]) the ring slot / visible count are
        read on the device, so the identical launch sequence can be replayed from a CUDA graph."""
        B, H, Dh, d, C = self.B, self.H, self.Dh, self.d, self.C
        ws = self.ws
        cb = int(self.bf16)
        tc = self.tc_linear
        wb = self.wsb if tc else None
        nv.call("commu_embed_fwd", tokens, self.emb32, d, d, math.sqrt(d), B, ws["x"], d, wb["x"] if tc else None, d)
        x = ws["x"]
        for l in range(self.L):
            w = self.W[l]
            if tc:
                nv.gemm(wb["x"], w["qkv"], m=B, n=3 * H * Dh, k=d, out_f32=ws["qkv"])
            else:
                self._linear(x, w["qkv"], None, False, None, ws["qkv"], B, 3 * H * Dh, d)
            nv.call("commu_pad_heads", ws["qkv"], 3 * H * Dh, 0, B, H, Dh, ws["q"], 0, H * 64, 64, 0, None)
            nv.call("commu_pad_heads", ws["qkv"], 3 * H * Dh, H * Dh, B, H, Dh, self.kc[l], cb, H * C * 64, C * 64,
                    slot * 64, dstate)
            nv.call("commu_pad_heads", ws["qkv"], 3 * H * Dh, 2 * H * Dh, B, H, Dh, self.vc[l], cb, H * C * 64, C * 64,
                    slot * 64, dstate)
            nv.call("commu_decode_attn", ws["q"], self.kc[l], self.vc[l], self.rt[l], cb, self.u, self.vb, B, H, C,
                    n_vis, slot, self.scale, ws["att"], H * 64, dstate, wb["att"] if tc else None)
            if tc:
                nv.gemm(wb["att"], w["o"], m=B, n=d, k=H * 64, add_f32=x, out_f32=ws["z"])
                nv.call("commu_layernorm_fwd", ws["z"], d, w["g1"], w["be1"], d, d, 1e-5, B, ws["y"], d, wb["y"], d,
                        None, None)
                nv.gemm(wb["y"], w["w1"], m=B, n=self.Di, k=d, bias=w["b1"], relu=True, out_bf16=wb["h"])
                nv.gemm(wb["h"], w["w2"], m=B, n=d, k=self.Di, bias=w["b2"], add_f32=ws["y"], out_f32=ws["z"])
                nv.call("commu_layernorm_fwd", ws["z"], d, w["g2"], w["be2"], d, d, 1e-5, B, ws["x"], d, wb["x"], d,
                        None, None)
            else:
                self._linear(ws["att"], w["o"], None, False, x, ws["z"], B, d, H * 64)
                self._ln(ws["z"], w["g1"], w["be1"], ws["y"])
                self._linear(ws["y"], w["w1"], w["b1"], True, None, ws["h"], B, self.Di, d)
                self._linear(ws["h"], w["w2"], w["b2"], False, ws["y"], ws["z"], B, d, self.Di)
                self._ln(ws["z"], w["g2"], w["be2"], ws["x"])
            x = ws["x"]
        if tc:
            nv.gemm(wb["x"], self.emb, m=B, n=self.V, k=d, bias=self.lbias, out_f32=ws["logits"])
        else:
            self._linear(x, self.emb, self.lbias, False, None, ws["logits"], B, self.V, d)

    @torch.no_grad()
    def prefill(self, ctx, state=None):
        """ctx: int64 [T, B]; feeds the tokens one at a time (same attention pattern as the reference's
        multi-token context call, model.py:606-628 with M = 0)."""
        state = state or DecodeState()
        logits = None
        for t in range(ctx.shape[0]):
            logits, state = self.step(ctx[t].contiguous(), state)
        return logits, state

    @torch.no_grad()
    def sample(self, logits, temperature, top_k=0, top_p=0.0, wrong=None, seed=0, offset=0, want_probs=False):
        """On-device sampler.  Returns (tokens int64 [B], probs [B,V] or None)."""
        B = logits.shape[0]
        toks = torch.empty(B, dtype=torch.int64, device=self.dev)
        probs = torch.empty(B, self.V, device=self.dev) if want_probs else None
        nv.call("commu_sample", logits, logits.stride(0), B, self.V, float(temperature), int(top_k), float(top_p),
                wrong, int(seed), int(offset), toks, probs, self.V, None)
        return toks, probs

    @torch.no_grad()
    def generate(self, ctx, n_new, temperature=0.95, top_k=0, top_p=0.9, seed=0, use_graph=True):
        """Batched generation (BASELINE config 4): ctx int64 [T0, B] -> tokens int64 [n_new, B].
        With use_graph the per-token launch sequence (embed, L layers, logits, sampler) is captured ONCE in
        a CUDA graph; ring slot, visible-key count and sampler counter live in a device int32[4]."""
        state = DecodeState()
        if ctx.shape[0] > 1:
            _, state = self.prefill(ctx[:-1], state)
        cur = ctx[-1].contiguous().clone()
        out = torch.empty(n_new, self.B, dtype=torch.int64, device=self.dev)
        if not use_graph:
            for t in range(n_new):
                logits, state = self.step(cur, state)
                cur, _ = self.sample(logits, temperature, top_k, top_p, None, seed, t)
                out[t] = cur
            return out
        dstate = torch.tensor([state.slot, 0, state.count, 0], dtype=torch.int32, device=self.dev)
        extra = 0 if self.same_length else 1

        def one_step():
            nv.call("commu_decode_advance", dstate, self.C, self.mem_len, extra)
            self._step_kernels(cur, 0, 1, dstate)
            nv.call("commu_sample", self.ws["logits"], self.V, self.B, self.V, float(temperature), int(top_k),
                    float(top_p), None, int(seed), 0, cur, None, self.V, dstate)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):        # warm-up outside capture (one real step)
            one_step()
            out[0] = cur
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step()
        for t in range(1, n_new):
            graph.replay()
            out[t] = cur
        self.last_state = dstate
        return out
