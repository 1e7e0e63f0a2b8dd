"""Native incremental decode engine: per-layer projected K/V ring cache + cached R-by-distance table,
one kernel sequence per generated token for a whole batch of sequences, on-device sampling.

Replaces the per-token `forward_generate` + `calc_probs` / `apply_sampling` / `infer_token` loop of
the reference (commu/midi_generator/midi_inferrer.py:199-237 over commu/model/model.py:606-628).
`precision="fp32"` streams fp32 weights / cache (greedy tokens bit-faithful to the fp32 reference up
to summation order); `precision="bf16"` halves the streamed bytes (throughput configuration).
"""
import math
import os

import torch

from commu import _native as nv


# ------------------------------------------------------------------------------------------------
# Index arithmetic of the TMA decode-attention kernel (csrc/decode_fused.cu, dec_attn_tma_kernel), restated on the host:
# documentation of the ring / table layout and the subject of tests/test_decode_index_cpu.py.  The kernel itself never
# calls these.
# ------------------------------------------------------------------------------------------------
def visible_slot_tiles(C, cur_slot, n_vis, tile=64):
    """Ring-slot tiles (index t covers slots t*tile .. t*tile+tile-1) that hold at least one of the n_vis newest
    entries (ages 0..n_vis-1, age 0 at cur_slot), in the order the kernel walks them.  C is a multiple of `tile`."""
    nts = C // tile
    lo = cur_slot - n_vis + 1
    if lo < 0:
        lo += C
    t_first = lo // tile
    nt = min(nts, ((lo % tile) + n_vis + tile - 1) // tile)
    return [(t_first + k) % nts for k in range(nt)]


def slot_age(C, cur_slot, slot):
    """Age of a ring slot (0 = newest); a slot is attended iff its age < n_vis."""
    return (cur_slot - slot) % C


def rtab2_row(C, cur_slot, slot):
    """Row of the reversed, doubled table rt2[h] (2C rows, rt2[j] = R[C-1 - (j mod C)]) that pairs with `slot`:
    consecutive slots of a tile map to consecutive rows starting at this value for the tile's first slot."""
    return (C - 1 - cur_slot + slot) % C


class DecodeState:
    """Immutable handle playing the role of the reference's `mems` in the decode loop: how many
    tokens are cached and where the newest sits in the ring.  Passing an older handle back rewinds."""
    __slots__ = ("count", "slot")

    def __init__(self, count=0, slot=-1):
        self.count, self.slot = count, slot


class DecodeEngine:
    lin_tiled = True            # fp32 / kernel-per-op linears on the register-tiled GEMM (COMMU_DECODE_LINEAR=simple: old kernel)
    _lin_scratch = _lin_cnt = None
    attn_splits, _att_part, _att_cnt = 1, None, None

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
        self.lin_tiled = os.environ.get("COMMU_DECODE_LINEAR", "tiled") == "tiled"
        # key splits of the streaming attention of the kernel-per-op engines: B*H*splits CTAs should be a near-integer
        # number of waves of the 2 x SM-count resident slots
        self.attn_splits = int(os.environ.get("COMMU_DECODE_ATTN_SPLITS", "0")) or self._pick_attn_splits()
        self._att_part = torch.empty(batch * self.H * self.attn_splits * 66, device=self.dev)
        self._att_cnt = torch.zeros(batch * self.H, dtype=torch.int32, device=self.dev)
        self._prepare()

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _prepare(self):
        dev, m = self.dev, self.m
        cdt = torch.bfloat16 if self.bf16 else torch.float32
        # bf16 throughput mode runs the fused token-step kernels (csrc/decode_fused.cu): 5 launches per layer.
        # attention kernel: 8 = stream-K TMA tensor-core kernel (persistent grid, `splits` CTAs per SM; ring capacity
        # rounded up to a multiple of 64 slots; +2 = 3-stage ring), 0 = SIMT cross-check kernel (`splits` key splits)
        self.fused = (self.bf16 and self.d <= 1024 and os.environ.get("COMMU_DECODE_FUSED", "1") != "0")
        self.attn_impl = int(os.environ.get("COMMU_DECODE_ATTN", "8"))
        if self.fused and self.attn_impl & 8:
            self.C = (self.mem_len + 1 + 63) // 64 * 64
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
        self.ws = dict(x=f(B, d), qkv=f(B, 3 * H * Dh), q=torch.zeros(B, H, 64, device=dev), att=f(B, H * 64), z=f(B, d), y=f(B, d),
                       h=f(B, self.Di), logits=f(B, self.V))
        # bf16 throughput mode: the linear layers run on the tcgen05 GEMM (weights streamed once per step,
        # the 64 batch rows are one half-filled 128-row MMA tile), so activations also exist as bf16 operands
        self.tc_linear = self.bf16 and d % 8 == 0 and self.Di % 8 == 0 and (H * Dh) % 8 == 0
        if self.fused:
            self._prepare_fused()
            self.tc_linear = False
        if self.tc_linear:
            hb = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
            self.wsb = dict(x=hb(B, d), att=hb(B, H * 64), y=hb(B, d), h=hb(B, self.Di))

    @torch.no_grad()
    def _prepare_fused(self):
        """Padded bf16 weights and the per-launch argument structs of the fused path.  Layouts: reduction
        lengths padded to multiples of 64 (zero columns), weight rows to multiples of 16 (zero rows), the qkv
        rows re-ordered into the padded head layout [q|k|v][H][64] so the kernel scatters without index maps."""
        dev, m = self.dev, self.m
        B, H, Dh, d, Di, C, V = self.B, self.H, self.Dh, self.d, self.Di, self.C, self.V
        r64 = lambda n: (n + 63) // 64 * 64
        r16 = lambda n: (n + 15) // 16 * 16
        dp, dip, hd = r64(d), r64(Di), H * 64
        self.pdl = int(os.environ.get("COMMU_DECODE_PDL", "1") != "0")
        self.splits = int(os.environ.get("COMMU_DECODE_SPLITS", "0")) or (2 if self.attn_impl & 8 else self._pick_splits())
        if self.attn_impl & 8:
            # reversed, doubled table per head [H, 2C, 64]: row j = R[C-1 - (j mod C)], so the 64 slots of a ring tile
            # are 64 consecutive rows even where the ages wrap (one TMA box)
            self.rt_h = [self._reverse_double(r) for r in self.rt]
        else:
            self.rt_h = self.rt
        ks = None
        for limit in (512, 1024):        # K chunk per CTA of the FF output GEMM (cluster split of d_inner)
            for cand in (1, 2, 4, 8):
                if ks is None and dip % (cand * 32) == 0 and dip // cand <= limit:
                    ks = cand
        if ks is None:
            raise RuntimeError("commu_b200 decode: d_inner %d is beyond the fused decode kernels (<= 8192)" % Di)
        sd = {n: p for n, p in m.named_parameters()}
        bf = torch.bfloat16
        zb = lambda *s: torch.zeros(*s, device=dev, dtype=bf)
        zf = lambda *s: torch.zeros(*s, device=dev)
        ws = self.wsf = dict(x=zf(B, d), y=zf(B, d), z1=zf(B, d), z2=zf(B, d), q=zf(B, H, 64), att=zb(B, hd),
                             h=zb(B, dip), part=zf(B * H * max(self.splits, (C + 63) // 64) * 66),
                             cnt=torch.zeros(B * H, dtype=torch.int32, device=dev))
        self.tok_buf = torch.zeros(B, dtype=torch.int64, device=dev)   # placeholder; step() points at its tokens
        self.fw, self.fargs = [], []
        mk = nv.dec_linear_args
        for l in range(self.L):
            pre = "layers.%d." % l
            wq = sd[pre + "dec_attn.qkv_net.weight"].detach()
            wqkv = zb(3 * hd, dp)
            wqkv.view(3, H, 64, dp)[:, :, :Dh, :d].copy_(wq.view(3, H, Dh, d))
            wo = zb(r16(d), hd)
            wo.view(r16(d), H, 64)[:d, :, :Dh].copy_(sd[pre + "dec_attn.o_net.weight"].detach().view(d, H, Dh))
            w1 = zb(dip, dp)
            w1[:Di, :d].copy_(sd[pre + "pos_ff.CoreNet.0.weight"].detach())
            b1 = zf(dip)
            b1[:Di].copy_(sd[pre + "pos_ff.CoreNet.0.bias"].detach())
            w2 = zb(r16(d), dip)
            w2[:d, :Di].copy_(sd[pre + "pos_ff.CoreNet.3.weight"].detach())
            W = self.W[l]
            self.fw.append((wqkv, wo, w1, b1, w2))
            prev = self.W[l - 1] if l else None
            common = dict(B=B, pdl=self.pdl, split_k=1)
            if l == 0:
                a_qkv = mk(prologue=nv.PRO_EMBED, tokens=self.tok_buf, emb=self.emb32, emb_scale=math.sqrt(d), **common)
            else:
                a_qkv = mk(prologue=nv.PRO_LN, z=ws["z2"], ldz=d, gamma=prev["g2"], beta=prev["be2"], eps=1e-5, **common)
            a_qkv.epilogue, a_qkv.K, a_qkv.N, a_qkv.d_true = nv.EPI_QKV, dp, 3 * hd, d
            a_qkv.x_out, a_qkv.ldx = ws["x"].data_ptr(), d
            a_qkv.w, a_qkv.ldw = wqkv.data_ptr(), dp
            a_qkv.q_out, a_qkv.k_cache, a_qkv.v_cache = ws["q"].data_ptr(), self.kc[l].data_ptr(), self.vc[l].data_ptr()
            a_qkv.H, a_qkv.C = H, C
            a_o = mk(prologue=nv.PRO_BF16, epilogue=nv.EPI_RES, K=hd, N=d, a_bf16=ws["att"], lda=hd, w=wo, ldw=hd,
                     res=ws["x"], ldr=d, out_f32=ws["z1"], ldo=d, **common)
            a_f1 = mk(prologue=nv.PRO_LN, epilogue=nv.EPI_RELU, K=dp, N=dip, d_true=d, z=ws["z1"], ldz=d, gamma=W["g1"],
                      beta=W["be1"], eps=1e-5, x_out=ws["y"], ldx=d, w=w1, ldw=dp, bias=b1, out_bf16=ws["h"], ldob=dip,
                      **common)
            a_f2 = mk(prologue=nv.PRO_BF16, epilogue=nv.EPI_RES, K=dip, N=d, a_bf16=ws["h"], lda=dip, w=w2, ldw=dip,
                      bias=W["b2"], res=ws["y"], ldr=d, out_f32=ws["z2"], ldo=d, B=B, pdl=self.pdl, split_k=ks)
            self.fargs.append((a_qkv, a_o, a_f1, a_f2))
        wl = zb(r16(V), dp)
        wl[:V, :d].copy_(self.emb32)
        last = self.W[-1]
        self.fw.append((wl,))
        self.a_logits = mk(prologue=nv.PRO_LN, epilogue=nv.EPI_LOGITS, K=dp, N=V, d_true=d, z=ws["z2"], ldz=d,
                           gamma=last["g2"], beta=last["be2"], eps=1e-5, w=wl, ldw=dp, bias=self.lbias,
                           out_f32=self.ws["logits"], ldo=V, **dict(B=B, pdl=self.pdl, split_k=1))

    @staticmethod
    def _reverse_double(r):
        """R by distance [C, H, 64] -> per head reversed and doubled [H, 2C, 64]: row j = R[C-1 - (j mod C)]."""
        return torch.cat([r.flip(0).permute(1, 0, 2)] * 2, dim=1).contiguous()

    def _pick_splits(self):
        """Key splits of the SIMT decode attention: B*H*splits CTAs should fill the SMs' resident slots (3 CTAs per
        SM) a near-integer number of times, with enough CTAs to even out the tail."""
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count if self.dev.type == "cuda" else 148
        slots, work = 3 * sms, self.B * self.H
        best, best_eff = 1, 0.0
        for s in range(1, 17):
            if self.mem_len // s < 128:
                break
            waves = work * s / slots
            eff = waves / math.ceil(waves)
            if eff > best_eff + 0.02:
                best, best_eff = s, eff
        return best

    def _attn_fused(self, l, slot, n_vis, dstate):
        """Layer l's single-query attention over its ring cache (q staged by the qkv launch) -> bf16 rows for o_net."""
        ws = self.wsf
        nv.call("commu_decode_attn_split", ws["q"], self.kc[l], self.vc[l], self.rt_h[l], self.u, self.vb, self.B, self.H,
                self.C, n_vis, slot, self.scale, self.splits, ws["part"], ws["cnt"], ws["att"], None, self.H * 64, dstate,
                self.pdl, self.attn_impl)

    def _step_fused(self, tokens, slot, n_vis, dstate):
        B, H, C = self.B, self.H, self.C
        ws = self.wsf
        if tokens.dtype != torch.int64 or not tokens.is_contiguous() or tokens.numel() != B:
            raise RuntimeError("commu_b200 decode: tokens must be a contiguous int64 [B] device tensor")
        self.fargs[0][0].tokens = tokens.data_ptr()
        dptr = dstate.data_ptr() if dstate is not None else None
        for l in range(self.L):
            a_qkv, a_o, a_f1, a_f2 = self.fargs[l]
            a_qkv.slot, a_qkv.dev_state = slot, dptr
            nv.dec_linear(a_qkv)
            self._attn_fused(l, slot, n_vis, dstate)
            nv.dec_linear(a_o)
            nv.dec_linear(a_f1)
            nv.dec_linear(a_f2)
        nv.dec_linear(self.a_logits)

    def _pick_attn_splits(self):
        sms = torch.cuda.get_device_properties(self.dev).multi_processor_count if self.dev.type == "cuda" else 148
        slots, work = 2 * sms, self.B * self.H
        best, best_eff = 1, 0.0
        for s_ in range(1, 9):
            if s_ > 1 and self.mem_len // s_ < 256:
                break
            waves = work * s_ / slots
            eff = waves / math.ceil(waves)
            if eff > best_eff + 0.03:
                best, best_eff = s_, eff
        return best

    def _lin_buffers(self, N):
        """Scratch of the K-split partial sums and the per-tile counters of the tiled linear kernel."""
        tiles = (N + 31) // 32
        need = 2 * 148 * 2048 + tiles * 2048
        if self._lin_scratch is None or self._lin_scratch.numel() < need or self._lin_cnt.numel() < tiles:
            self._lin_scratch = torch.empty(max(need, 4 << 20), device=self.dev)
            self._lin_cnt = torch.zeros(max(tiles, 1024), dtype=torch.int32, device=self.dev)

    def _linear(self, x, w, bias, relu, res, out, B, N, K):
        if self.lin_tiled:
            # register-tiled SIMT GEMM with K splits (fixed-order split reduction: run-to-run identical results)
            self._lin_buffers(N)
            nv.call("commu_decode_linear_tiled", x, x.stride(0), w, w.stride(0), int(w.dtype == torch.bfloat16), bias,
                    int(relu), res, res.stride(0) if res is not None else 0, out, out.stride(0), B, N, K, 0,
                    self._lin_scratch, self._lin_cnt)
            return
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
        if self.fused:
            return self._step_fused(tokens, slot, n_vis, dstate)
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
            elif self.lin_tiled and not self.bf16:
                # fp32 engine: the qkv projection scatters q / k / v into the padded head layouts from its epilogue
                self._lin_buffers(3 * H * Dh)
                nv.call("commu_decode_qkv_tiled", x, x.stride(0), w["qkv"], w["qkv"].stride(0), B, H, Dh, d, ws["q"],
                        self.kc[l], self.vc[l], C, slot, dstate, self._lin_scratch, self._lin_cnt)
            else:
                self._linear(x, w["qkv"], None, False, None, ws["qkv"], B, 3 * H * Dh, d)
            if tc or self.bf16 or not self.lin_tiled:
                nv.call("commu_pad_heads", ws["qkv"], 3 * H * Dh, 0, B, H, Dh, ws["q"], 0, H * 64, 64, 0, None)
                nv.call("commu_pad_heads", ws["qkv"], 3 * H * Dh, H * Dh, B, H, Dh, self.kc[l], cb, H * C * 64, C * 64,
                        slot * 64, dstate)
                nv.call("commu_pad_heads", ws["qkv"], 3 * H * Dh, 2 * H * Dh, B, H, Dh, self.vc[l], cb, H * C * 64, C * 64,
                        slot * 64, dstate)
            nv.call("commu_decode_attn", ws["q"], self.kc[l], self.vc[l], self.rt[l], cb, self.u, self.vb, B, H, C,
                    n_vis, slot, self.scale, ws["att"], H * 64, dstate, wb["att"] if tc else None,
                    self.attn_splits, self._att_part, self._att_cnt)
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
                cur, _ = self.sample(logits, temperature, top_k, top_p, None, seed, t + 1)   # counter t+1: same draws as the graph path
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
