"""CPU dry-run of the host-side call sequence: the native layer is replaced by a stub that only
validates arity / argument kinds against the typed signature table, so Python plumbing errors
(wrong argument counts, shapes, None handling) surface without a GPU.  No arithmetic is checked."""
from types import SimpleNamespace as NS

import pytest
import torch

from commu import _native as nv


class _Stub:
    def __init__(self):
        self.calls = []

    def call(self, name, *args):
        sig = nv.SIGNATURES[name]
        assert len(args) in (len(sig), len(sig) - 1), (name, len(args), len(sig))
        for a, t in zip(args, sig):
            if t is nv.P:
                assert a is None or hasattr(a, "data_ptr") or isinstance(a, (bytes, int)), (name, type(a))
            elif t in (nv.I, nv.L, nv.U):
                assert isinstance(a, (int, bool)), (name, a, type(a))
            elif t is nv.F:
                assert isinstance(a, (int, float)), (name, a)
        self.calls.append(name)

    def dec_linear(self, a):
        assert isinstance(a, nv.DecLinearArgs)
        assert a.prologue in (nv.PRO_EMBED, nv.PRO_LN, nv.PRO_BF16) and a.epilogue in range(4)
        assert a.w and a.K % 64 == 0 and a.N > 0 and 1 <= a.B <= 64 and a.ldw == a.K
        if a.prologue == nv.PRO_LN:
            assert a.z and a.gamma and a.beta and a.eps > 0 and 0 < a.d_true <= a.K
        if a.prologue == nv.PRO_EMBED:
            assert a.tokens and a.emb and a.emb_scale > 0
        if a.prologue == nv.PRO_BF16:
            assert a.a_bf16 and a.lda >= a.K and a.K % (a.split_k * 32) == 0
        if a.epilogue == nv.EPI_QKV:
            assert a.q_out and a.k_cache and a.v_cache and a.N == 3 * a.H * 64 and 0 <= a.slot < a.C
        if a.epilogue == nv.EPI_RES:
            assert a.res and a.out_f32
        if a.epilogue == nv.EPI_RELU:
            assert a.out_bf16 and a.N % 16 == 0
        self.calls.append("dec_linear")

    def gemm(self, a, b, **kw):
        assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16
        for k in ("m", "n", "k"):
            assert kw[k] > 0
        self.calls.append("gemm")


@pytest.fixture
def stub(monkeypatch):
    s = _Stub()
    monkeypatch.setattr(nv, "call", s.call)
    monkeypatch.setattr(nv, "gemm", s.gemm)
    monkeypatch.setattr(nv, "dec_linear", s.dec_linear)
    monkeypatch.setattr(nv, "lib", lambda: None)
    import commu.engine.native_lm as nl
    import commu.engine.decode as dec
    monkeypatch.setattr(nl.NativeLM, "__init__", _patched_init(nl.NativeLM.__init__))
    monkeypatch.setattr(dec.DecodeEngine, "__init__", _patched_init(dec.DecodeEngine.__init__))
    return s


def _patched_init(orig):
    def init(self, *a, **k):
        class FakeDev:
            type = "cuda"
        # run the original with the device check neutralised
        import torch as _t
        real = _t.Tensor.device
        try:
            return orig(self, *a, **k)
        except RuntimeError as e:
            if "CUDA" not in str(e) and "cuda" not in str(e):
                raise
            raise
    return init


def _model(d=60, H=6, Di=100, L=2, V=53, T=8, M=24, same=True):
    from commu.model.model import MemTransformerLM

    class Vc:
        def __len__(self):
            return V
    cfg = NS(MODEL=NS(num_layers=L, num_heads=H, units=d, inner_size=Di, dropout=0.0, attention_dropout=0.0,
                      same_length=same, clamp_len=-1), TRAIN=NS(tgt_length=T, mem_length=M))
    return MemTransformerLM(cfg, Vc())


def test_train_path_plumbing(stub, monkeypatch):
    import commu.engine.native_lm as nl
    m = _model()
    # bypass the CUDA-only guards for the dry run
    monkeypatch.setattr(nl.NativeLM, "__init__", _cpu_init(nl.NativeLM))
    monkeypatch.setattr(type(m), "_check_inputs", lambda self, d, mm: None)
    data = torch.randint(1, 53, (8, 2))
    target = torch.randint(0, 53, (8, 2))
    mems = None
    for s in range(3):
        loss, mems = m(data, target, torch.tensor([False, s == 1]), mems)
        assert loss.shape == (8, 2) and tuple(mems.shape) == (3, min(24, 8 * (s + 1)), 2, 60)
        loss.mean().backward()
    assert all(p.grad is not None for n, p in m.named_parameters())
    assert "commu_relattn_bwd" in stub.calls and "commu_embed_bwd" in stub.calls
    lg, mems = m.forward_generate(data[:1], mems)
    assert lg.shape == (1, 2, 53)


def _cpu_init(cls):
    def init(self, params, n_layer, n_head, d_model, d_inner, n_token, inv_freq):
        self.P = params
        self.L, self.H, self.d, self.Di, self.V = n_layer, n_head, d_model, d_inner, n_token
        self.Dh = d_model // n_head
        c = lambda a: (a + 63) // 64 * 64
        self.dp, self.dip, self.vp = c(d_model), c(d_inner), c(n_token)
        self.hd = n_head * 64
        self.inv_freq = inv_freq
        self.dev = params["r_w_bias"].device
        self.aligned = False
        self._shadow = None
        self._shadow_version = None
        self._pos_cache = {}
        self.saved = None
        self.attn_fwd_impl = "commu_relattn_fwd"
        self.bwd_materialise = False
    return init


def test_trainer_plumbing(stub, monkeypatch):
    import commu.engine.native_lm as nl
    from commu.engine.trainer import Trainer
    monkeypatch.setattr(nl.NativeLM, "__init__", _cpu_init(nl.NativeLM))
    m = _model(d=64, H=1, Di=128, same=False)
    tr = Trainer(m, lr=0.004, warmup_step=2, lr_min=1e-4, batch_chunk=2)
    data = torch.randint(1, 53, (8, 4))
    target = torch.randint(0, 53, (8, 4))
    for s in range(2):
        loss, gn = tr.train_step(data, target, torch.tensor([False, True, False, False]))
        assert loss.dim() == 0
    sd = tr.optimizer_state_dict()
    assert len(sd["state"]) == len(m._param_list)
    tr.load_optimizer_state_dict(sd)
    assert "commu_clip_adam" in stub.calls and "commu_sumsq" in stub.calls


def test_exchange_spans_partition_the_gradient_arena(stub, monkeypatch):
    """The overlapped gradient exchange sends one span per decoder layer plus the remainder (biases, embedding,
    logits bias): together they must cover the flat arena exactly once, and a layer's span must hold exactly that
    layer's parameters (train.py:155, 467-473 of the reference all-reduces every gradient once per step)."""
    import commu.engine.native_lm as nl
    from commu.engine.trainer import Trainer
    monkeypatch.setattr(nl.NativeLM, "__init__", _cpu_init(nl.NativeLM))
    m = _model(d=64, H=1, Di=128, L=3, same=False)
    tr = Trainer(m, lr=0.004)
    total = tr.flat_g.numel()
    spans = sorted(tr.layer_spans + tr.other_spans)
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    for l, (lo, hi) in enumerate(tr.layer_spans):
        inside = [n for n, o in tr.offsets.items() if lo <= o < hi]
        assert inside and all(n.startswith("layers.%d." % l) for n in inside)
        assert len(inside) == len([n for n in tr.offsets if n.startswith("layers.%d." % l)])


def test_lr_schedule_floor_is_world_independent(stub, monkeypatch):
    """train.py:441-461 of the reference: the optimizer runs at cfg.TRAIN.lr / num_gpus, but the floor of the
    inverse-sqrt schedule is the ratio lr_min / cfg.TRAIN.lr with the UNDIVIDED lr.  At 8 ranks the floor multiplier
    must stay 1e-4 / 0.004 = 0.025 (not 0.2), i.e. lr keeps decaying past step 2500 down to lr_min / 8."""
    import commu.engine.native_lm as nl
    from commu.engine.trainer import Trainer, lr_multiplier
    monkeypatch.setattr(nl.NativeLM, "__init__", _cpu_init(nl.NativeLM))
    lr, lr_min, warm = 0.004, 1e-4, 100
    for world in (1, 2, 8):
        m = _model(d=64, H=1, Di=128, same=False)
        tr = Trainer(m, lr=lr / world, warmup_step=warm, lr_min=lr_min, world=world, global_lr=lr)
        for step in (0, 1, 50, 100, 101, 2500, 20000, 10 ** 6):
            tr.step = step
            ref = (lr / world) * lr_multiplier(step, warm, lr, lr_min)        # LambdaLR(lr_lambda) * local_lr
            assert abs(tr.current_lr() - ref) < 1e-15, (world, step)
        tr.step = 20000
        assert abs(tr.current_lr() - (lr / world) * (warm ** 0.5) / (20000 ** 0.5)) < 1e-12   # still decaying
        tr.step = 10 ** 7
        assert abs(tr.current_lr() - lr_min / world) < 1e-15                                   # the floor


def test_decode_plumbing(stub, monkeypatch):
    import commu.engine.decode as dec
    m = _model()
    orig = dec.DecodeEngine.__init__

    def init(self, model, batch, mem_len, same_length=True, precision="fp32"):
        class D:
            type = "cuda"
        real_dev = model.r_w_bias.device
        # neutralise only the device-type check
        monkeypatch.setattr(dec.torch.Tensor, "device", property(lambda s: real_dev), raising=False)
        self.m = model
        self.B, self.mem_len, self.same_length = batch, mem_len, bool(same_length)
        self.bf16 = precision == "bf16"
        self.L, self.H, self.d, self.Dh = model.n_layer, model.n_head, model.d_model, model.d_head
        self.Di, self.V = model.d_inner, model.n_token
        self.C = mem_len + 1
        self.dev = real_dev
        self.scale = 1.0
        self._prepare()
    monkeypatch.setattr(dec.DecodeEngine, "__init__", init)
    for prec in ("fp32", "bf16"):
        eng = dec.DecodeEngine(m, 3, 24, True, prec)
        ctx = torch.randint(1, 53, (5, 3))
        out = eng.generate(ctx, 4, temperature=0.95, top_k=0, top_p=0.9, seed=1, use_graph=False)
        assert out.shape == (4, 3)
    assert "commu_decode_attn" in stub.calls and "commu_sample" in stub.calls
    assert "commu_decode_attn_split" in stub.calls and "dec_linear" in stub.calls      # bf16 = fused step
    monkeypatch.setenv("COMMU_DECODE_FUSED", "0")
    n0 = stub.calls.count("gemm")
    eng = dec.DecodeEngine(m, 3, 24, True, "bf16")
    eng.generate(torch.randint(1, 53, (5, 3)), 2, temperature=0.95, top_k=0, top_p=0.9, seed=1, use_graph=False)
    assert not eng.fused


def test_fused_decode_weight_layouts(stub, monkeypatch):
    """The padded bf16 weights of the fused decode step (DecodeEngine._prepare_fused) reproduce the unpadded linear
    layers: q / k / v in the padded head layout [q|k|v][H][64], o_net over the padded head columns, FF matrices with
    zero padding, tied logits - checked with plain CPU matmuls (d = 60, Dh = 10, d_inner = 100: every padding bites)."""
    import commu.engine.decode as dec
    torch.manual_seed(0)
    m = _model()
    with torch.no_grad():
        for p in m.parameters():
            p.normal_(0.0, 0.5)

    def init(self, model, batch, mem_len, same_length=True, precision="fp32"):
        real_dev = model.r_w_bias.device
        monkeypatch.setattr(dec.torch.Tensor, "device", property(lambda s: real_dev), raising=False)
        self.m = model
        self.B, self.mem_len, self.same_length = batch, mem_len, bool(same_length)
        self.bf16 = precision == "bf16"
        self.L, self.H, self.d, self.Dh = model.n_layer, model.n_head, model.d_model, model.d_head
        self.Di, self.V = model.d_inner, model.n_token
        self.C = mem_len + 1
        self.dev = real_dev
        self.scale = 1.0
        self._prepare()
    monkeypatch.setattr(dec.DecodeEngine, "__init__", init)
    eng = dec.DecodeEngine(m, 3, 24, True, "bf16")
    assert eng.fused and eng.C % 64 == 0 and eng.C >= 25
    H, Dh, d, Di, V = eng.H, eng.Dh, eng.d, eng.Di, eng.V
    sd = dict(m.named_parameters())
    x = torch.randn(3, d)
    xp = torch.zeros(3, 64)
    xp[:, :d] = x
    tol = dict(rtol=2e-2, atol=2e-2)            # the padded copies are bf16
    for l in range(eng.L):
        wqkv, wo, w1, b1, w2 = [t.float() for t in eng.fw[l]]
        pre = "layers.%d." % l
        ref = (x @ sd[pre + "dec_attn.qkv_net.weight"].t()).view(3, 3, H, Dh)
        got = (xp @ wqkv.t()).view(3, 3, H, 64)
        assert torch.allclose(got[..., :Dh], ref, **tol) and float(got[..., Dh:].abs().max()) == 0.0
        att = torch.randn(3, H, Dh)
        attp = torch.zeros(3, H, 64)
        attp[..., :Dh] = att
        ref_o = att.reshape(3, H * Dh) @ sd[pre + "dec_attn.o_net.weight"].t()
        assert torch.allclose((attp.reshape(3, H * 64) @ wo.t())[:, :d], ref_o, **tol)
        ref_h = x @ sd[pre + "pos_ff.CoreNet.0.weight"].t() + sd[pre + "pos_ff.CoreNet.0.bias"]
        got_h = xp @ w1.t() + b1
        assert torch.allclose(got_h[:, :Di], ref_h, **tol) and float(got_h[:, Di:].abs().max()) == 0.0
        hh = torch.zeros(3, w2.shape[1])
        hh[:, :Di] = torch.randn(3, Di)
        assert torch.allclose((hh @ w2.t())[:, :d], hh[:, :Di] @ sd[pre + "pos_ff.CoreNet.3.weight"].t(), **tol)
    wl = eng.fw[-1][0].float()
    assert torch.allclose((xp @ wl.t())[:, :V], x @ sd["word_emb.emb_layers.0.weight"].t(), **tol)
    # reversed, doubled relative-position table of the TMA kernel: rt2[h, j] = rt[C-1 - (j mod C), h]
    C = eng.C
    rt = torch.randn(C, H, 64)
    rt2 = dec.DecodeEngine._reverse_double(rt)
    assert rt2.shape == (H, 2 * C, 64) and eng.rt_h[0].shape == (H, 2 * C, 64)
    j = torch.arange(2 * C)
    assert torch.equal(rt2, rt[C - 1 - (j % C)].permute(1, 0, 2))
