"""Drop-in `commu.model.model`: same class name, constructor, parameter names / shapes and method
signatures as the reference (commu/model/model.py:423-693), but every forward / backward runs in
the hand-written sm_100a kernels of libcommu_b200.so through `commu.engine.native_lm.NativeLM`.

The sub-modules below only HOLD parameters so that `state_dict()` keys match reference checkpoints
(SURVEY.md section 8b); their own `forward` is never used.  There is no CPU / eager fallback: calling
the model on a non-CUDA device, or without the native library, raises.
"""
import logging

import torch
import torch.nn as nn

from commu.engine.native_lm import Mems, NativeLM

_log = logging.getLogger("ComMU")


class _Holder(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - containers only
        raise RuntimeError("parameter container; the computation lives in the native engine")


class _AttnParams(_Holder):
    def __init__(self, n_head, d_model, d_head):
        super().__init__()
        self.qkv_net = nn.Linear(d_model, 3 * n_head * d_head, bias=False)
        self.o_net = nn.Linear(n_head * d_head, d_model, bias=False)
        self.layer_norm = nn.LayerNorm(d_model)
        self.r_net = nn.Linear(d_model, n_head * d_head, bias=False)


class _FFParams(_Holder):
    def __init__(self, d_model, d_inner):
        super().__init__()
        # indices 0 and 3 carry weights, like the reference nn.Sequential (model.py:163-169)
        self.CoreNet = nn.ModuleList([nn.Linear(d_model, d_inner), nn.Identity(), nn.Identity(),
                                      nn.Linear(d_inner, d_model), nn.Identity()])
        self.layer_norm = nn.LayerNorm(d_model)


class _LayerParams(_Holder):
    def __init__(self, n_head, d_model, d_head, d_inner):
        super().__init__()
        self.dec_attn = _AttnParams(n_head, d_model, d_head)
        self.pos_ff = _FFParams(d_model, d_inner)


class _EmbParams(_Holder):
    def __init__(self, n_token, d_model):
        super().__init__()
        self.emb_layers = nn.ModuleList([nn.Embedding(n_token, d_model)])
        self.emb_projs = nn.ParameterList()


class _CritParams(_Holder):
    def __init__(self, n_token, d_model):
        super().__init__()
        self.out_layers = nn.ModuleList([nn.Linear(d_model, n_token)])
        self.out_projs = nn.ParameterList()
        self.n_clusters = 0


class _PosParams(_Holder):
    def __init__(self, d_model):
        super().__init__()
        self.register_buffer("inv_freq", 1.0 / (10000 ** (torch.arange(0.0, d_model, 2.0) / d_model)))


class _NativeLoss(torch.autograd.Function):
    """Whole-model autograd node: forward = native forward, backward = native backward."""

    @staticmethod
    def forward(ctx, model, data, target, reset, mems, *params):
        eng = model._engine()
        need = model._need_save   # grad mode is always off inside Function.forward; decided by the caller
        loss, new_mems = eng.forward_loss(data, target, reset, mems, model.mem_len, model.same_length,
                                          model.clamp_len, save=need, dropout=model._dropout_arg())
        model._last_mems = new_mems
        ctx.model = model
        ctx.saved_ctx = eng.saved
        eng.saved = None
        return loss

    @staticmethod
    def backward(ctx, dloss):
        model = ctx.model
        eng = model._engine()
        names = model._param_names
        grads = {n: torch.zeros_like(p) for n, p in zip(names, model._param_list)}
        eng.saved = ctx.saved_ctx
        eng.backward(dloss.contiguous(), grads)
        return (None, None, None, None, None) + tuple(grads[n] for n in names)


class MemTransformerLM(nn.Module):
    def __init__(self, cfg, vocab):
        super().__init__()
        m, t = cfg.MODEL, cfg.TRAIN
        self.cfg = cfg
        self.n_token = len(vocab)
        self.n_layer, self.n_head, self.d_model = m.num_layers, m.num_heads, m.units
        self.d_head = m.units // m.num_heads
        self.d_inner = m.inner_size
        self.d_embed = m.units
        self.dropout_p, self.dropatt_p = float(m.dropout), float(m.attention_dropout)
        self.tgt_len, self.mem_len = t.tgt_length, t.mem_length
        self.max_klen = self.tgt_len + self.mem_len
        self.same_length, self.clamp_len = m.same_length, m.clamp_len
        self.detach_mems_grad = True

        self.word_emb = _EmbParams(self.n_token, self.d_model)
        self.layers = nn.ModuleList([_LayerParams(self.n_head, self.d_model, self.d_head, self.d_inner)
                                     for _ in range(self.n_layer)])
        self.crit = _CritParams(self.n_token, self.d_model)
        self.crit.out_layers[0].weight = self.word_emb.emb_layers[0].weight   # tied (model.py:480-481)
        self.pos_emb = _PosParams(self.d_model)
        self.r_w_bias = nn.Parameter(torch.zeros(self.n_head, self.d_head))
        self.r_r_bias = nn.Parameter(torch.zeros(self.n_head, self.d_head))
        self._eng = None
        self._last_mems = None

    # ---- reference surface -------------------------------------------------------------------------
    def reset_length(self, tgt_len, mem_len):
        self.tgt_len, self.mem_len = tgt_len, mem_len

    def init_mems(self, n_layers=None):
        return None  # an empty memory is represented by None in the native engine

    def forward(self, data, target, reset_mems, mems):
        """-> (per-token NLL [T,B] fp32, new_mems)   (reference model.py:678-693)"""
        self._check_inputs(data, mems)
        plist = self._params()
        self._need_save = torch.is_grad_enabled() and any(p.requires_grad for p in plist)
        loss = _NativeLoss.apply(self, data, target, reset_mems, mems, *plist)
        new_mems, self._last_mems = self._last_mems, None
        return loss, new_mems

    def forward_generate(self, data, mems):
        """-> (logits [T,B,V] fp32, new_mems)   (reference model.py:606-628)"""
        self._check_inputs(data, mems)
        with torch.no_grad():
            return self._engine().forward_logits(data, mems, self.mem_len, self.same_length, self.clamp_len)

    # ---- plumbing ----------------------------------------------------------------------------------
    def _check_inputs(self, data, mems):
        if not data.is_cuda:
            raise RuntimeError("commu_b200: MemTransformerLM runs only on CUDA (sm_100a); there is no CPU path")
        if mems is not None and not isinstance(mems, Mems):
            if hasattr(mems, "numel") and mems.numel() == 0:
                return
            raise TypeError("commu_b200: `mems` must be None or the handle returned by a previous call")

    def _dropout_arg(self):
        """(p, p_att, seed) of this forward in training mode (reference: nn.Dropout modules, model.py:166-168,
        210-211, 454), else None.  Every forward draws a fresh 63-bit seed from torch's CPU generator, so
        `torch.manual_seed` makes runs repeatable; masks themselves come from the kernels' counter-based RNG."""
        if not self.training or (self.dropout_p <= 0 and self.dropatt_p <= 0):
            return None
        seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        return (self.dropout_p, self.dropatt_p, seed)

    def _params(self):
        named = [(n, p) for n, p in self.named_parameters() if n != "crit.out_layers.0.weight"]
        self._param_names = [n for n, _ in named]
        self._param_list = [p for _, p in named]
        return self._param_list

    def _engine(self):
        dev = self.r_w_bias.device
        self._params()
        if self._eng is None or self._eng.dev != dev or self._eng_ids != [id(p) for p in self._param_list]:
            P = dict(zip(self._param_names, self._param_list))
            self._eng = NativeLM(P, self.n_layer, self.n_head, self.d_model, self.d_inner, self.n_token,
                                 self.pos_emb.inv_freq)
            self._eng_ids = [id(p) for p in self._param_list]
            self._eng_ptrs = [p.data_ptr() for p in self._param_list]
        elif self._eng_ptrs != [p.data_ptr() for p in self._param_list]:
            # parameters were re-materialised (load_state_dict keeps storage; .to() / flatten does not)
            self._eng.P = dict(zip(self._param_names, self._param_list))
            self._eng_ptrs = [p.data_ptr() for p in self._param_list]
            self._eng._shadow_version = None
        return self._eng
