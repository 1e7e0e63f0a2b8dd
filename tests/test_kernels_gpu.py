"""GPU parity of the non-GEMM kernels against plain torch fp32 math (same formulas as the oracle)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _relattn_ref(q, k, v, r, u, vb, reset, T, M, B, H, Dh, same_length, shift, scale):
    """q [T,B,H,64] k,v [K,B,H,64] r [K,H,64] fp32 (already bf16-rounded values)."""
    K = T + M
    qu = (q + u).bfloat16().float()
    qv = (q + vb).bfloat16().float()
    AC = torch.einsum("ibhd,jbhd->bhij", qu, k)
    QR = torch.einsum("ibhd,thd->bhit", qv, r)
    ii = torch.arange(T, device=q.device)[:, None]
    jj = torch.arange(K, device=q.device)[None, :]
    dist = (ii + M - jj).clamp(min=0)
    BD = QR.gather(3, dist[None, None].expand(B, H, T, K))
    s = (AC + BD) * scale
    ok = jj <= ii + M
    if same_length:
        ok = ok & (jj > ii - shift)
    ok = ok[None].expand(B, T, K).clone()
    if reset is not None:
        ok[reset.bool()] &= (jj >= M)
    s = s.masked_fill(~ok[:, None], float("-inf"))
    lse = torch.logsumexp(s, dim=-1)
    p = torch.softmax(s, dim=-1)
    out = torch.einsum("bhij,jbhd->ibhd", p, v)
    return out, lse, p, s


ATT_CASES = [
    # T, M, B, H, same_length, mem_len, with_reset
    (64, 0, 1, 1, 0, 64, 0),
    (64, 64, 2, 2, 0, 64, 1),
    (100, 37, 2, 3, 0, 128, 1),
    (128, 128, 1, 2, 1, 128, 0),
    (96, 160, 2, 2, 1, 200, 1),
    (1, 77, 3, 2, 1, 77, 0),
    (256, 256, 1, 1, 0, 256, 0),
]


@pytest.mark.parametrize("T,M,B,H,same_length,mem_len,with_reset", ATT_CASES + [(300, 500, 2, 2, 0, 512, 1),
                                                                         (384, 384, 1, 2, 1, 384, 0)])
@pytest.mark.parametrize("impl", ["commu_relattn_fwd", "commu_relattn_fwd_tc"])
def test_relattn_fwd(T, M, B, H, same_length, mem_len, with_reset, impl):
    from commu import _native as nv
    torch.manual_seed(T * 13 + M)
    dev = "cuda"
    K = T + M
    Dh = 64
    q = torch.randn(T, B, H, Dh, device=dev).bfloat16()
    kv = torch.randn(K, B, 2, H, Dh, device=dev).bfloat16()
    r = torch.randn(K, H, Dh, device=dev).bfloat16()
    u = torch.randn(H, Dh, device=dev) * 0.5
    vb = torch.randn(H, Dh, device=dev) * 0.5
    reset = (torch.rand(B, device=dev) < 0.5) if with_reset else None
    if with_reset:
        reset[0] = True
    mask_len = K - mem_len
    shift = T - mask_len if mask_len > 0 else T
    scale = 1.0 / math.sqrt(Dh)
    out = torch.zeros(T, B, H * Dh, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    k_t = kv[:, :, 0]
    v_t = kv[:, :, 1]
    reset_u8 = reset.to(torch.uint8) if reset is not None else None
    extra = (None, None) if impl == "commu_relattn_fwd_tc" else ()      # p_save, mt_save (tcgen05 entry point only)
    nv.call(impl, q, H * Dh, k_t, v_t, 2 * H * Dh, r, H * Dh, K, u, vb, reset_u8,
            T, M, B, H, same_length, shift, scale, out, H * Dh, lse, None, None, *extra)
    torch.cuda.synchronize()
    ref, ref_lse, _, _ = _relattn_ref(q.float(), k_t.float(), v_t.float(), r.float(), u, vb, reset, T, M,
                                      B, H, Dh, same_length, shift, scale)
    err = (out.view(T, B, H, Dh).float() - ref).abs().max().item()
    lerr = (lse - ref_lse).abs().max().item()
    assert err < 0.03, err          # bf16 P and bf16 output rounding
    assert lerr < 2e-3, lerr


@pytest.mark.parametrize("T,M,B,H,same_length,mem_len,with_reset", ATT_CASES + [(300, 500, 2, 2, 0, 512, 1),
                                                                         (384, 384, 1, 2, 1, 384, 0)])
@pytest.mark.parametrize("impl", ["mat", "tc", "v1"])
def test_relattn_bwd(T, M, B, H, same_length, mem_len, with_reset, impl):
    """mat = product path (probabilities stored by the tcgen05 forward, dS materialised once, band GEMMs);
    tc / v1 = the recompute passes (tcgen05 / warp-MMA)."""
    from commu import _native as nv
    L = nv.lib()
    if impl == "mat":
        _relattn_bwd_case(nv, T, M, B, H, same_length, mem_len, with_reset, fwd_impl="commu_relattn_fwd_tc", mat=True)
        return
    flag = 1 if impl == "tc" else 0
    L.commu_relattn_bwd_set_impl(flag, flag, flag)
    try:
        _relattn_bwd_case(nv, T, M, B, H, same_length, mem_len, with_reset)
    finally:
        L.commu_relattn_bwd_set_impl(1, 1, 1)


def _relattn_bwd_case(nv, T, M, B, H, same_length, mem_len, with_reset, fwd_impl="commu_relattn_fwd", mat=False):
    torch.manual_seed(T * 17 + M)
    dev = "cuda"
    K = T + M
    Dh = 64
    q = (torch.randn(T, B, H, Dh, device=dev) * 0.7).bfloat16()
    kv = (torch.randn(K, B, 2, H, Dh, device=dev) * 0.7).bfloat16()
    r = (torch.randn(K, H, Dh, device=dev) * 0.7).bfloat16()
    u = torch.randn(H, Dh, device=dev) * 0.5
    vb = torch.randn(H, Dh, device=dev) * 0.5
    reset = (torch.rand(B, device=dev) < 0.5) if with_reset else None
    if with_reset:
        reset[0] = True
    mask_len = K - mem_len
    shift = T - mask_len if mask_len > 0 else T
    scale = 1.0 / math.sqrt(Dh)
    out = torch.zeros(T, B, H * Dh, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    qu_s = torch.zeros(T, B, H * Dh, device=dev, dtype=torch.bfloat16)
    qv_s = torch.zeros_like(qu_s)
    k_t, v_t = kv[:, :, 0], kv[:, :, 1]
    reset_u8 = reset.to(torch.uint8) if reset is not None else None
    psv = mtv = ws = None
    extra = ()
    if fwd_impl == "commu_relattn_fwd_tc":
        if mat:
            p_bytes, mt_bytes, _, _ = nv.attn_sizes(T, M, B, H)
            # poisoned, not zeroed: the backward may only read what the forward wrote
            psv = torch.full((p_bytes // 2,), float("nan"), dtype=torch.bfloat16, device=dev)
            mtv = torch.full((mt_bytes // 4,), float("nan"), device=dev)
            ws = nv.attn_bwd_workspace(T, M, B, H, dev)
        extra = (psv, mtv)
    nv.call(fwd_impl, q, H * Dh, k_t, v_t, 2 * H * Dh, r, H * Dh, K, u, vb, reset_u8,
            T, M, B, H, same_length, shift, scale, out, H * Dh, lse, qu_s, qv_s, *extra)
    dout = (torch.randn(T, B, H * Dh, device=dev) * 0.5).bfloat16()
    delta = torch.empty(B, H, T, device=dev)
    dq = torch.zeros(T, B, H * Dh, device=dev, dtype=torch.bfloat16)
    dkv = torch.full((K, B, 2, H, Dh), 9.0, device=dev, dtype=torch.bfloat16)
    dr = torch.zeros(K, H * Dh, device=dev)
    du = torch.zeros(H, Dh, device=dev)
    dvb = torch.zeros(H, Dh, device=dev)
    nv.call("commu_relattn_bwd", qu_s, qv_s, H * Dh, k_t, v_t, 2 * H * Dh, r, H * Dh, K, reset_u8,
            T, M, B, H, same_length, shift, scale, out, H * Dh, lse, dout, H * Dh, delta,
            dq, H * Dh, dkv[:, :, 0], dkv[:, :, 1], 2 * H * Dh, dr, du, dvb, psv, mtv, ws,
            ws.numel() if ws is not None else 0)
    torch.cuda.synchronize()
    # autograd reference on the same bf16-rounded operands (qu, qv are leaves: d/dq = d/dqu + d/dqv)
    quf = qu_s.view(T, B, H, Dh).float().requires_grad_(True)
    qvf = qv_s.view(T, B, H, Dh).float().requires_grad_(True)
    kf = k_t.float().requires_grad_(True)
    vf = v_t.float().requires_grad_(True)
    rf = r.float().requires_grad_(True)
    AC = torch.einsum("ibhd,jbhd->bhij", quf, kf)
    QR = torch.einsum("ibhd,thd->bhit", qvf, rf)
    ii = torch.arange(T, device=dev)[:, None]
    jj = torch.arange(K, device=dev)[None, :]
    dist = (ii + M - jj).clamp(min=0)
    BD = QR.gather(3, dist[None, None].expand(B, H, T, K))
    s = (AC + BD) * scale
    ok = jj <= ii + M
    if same_length:
        ok = ok & (jj > ii - shift)
    ok = ok[None].expand(B, T, K).clone()
    if reset is not None:
        ok[reset.bool()] &= (jj >= M)
    s = s.masked_fill(~ok[:, None], float("-inf"))
    pr = torch.softmax(s, dim=-1)
    o_ref = torch.einsum("bhij,jbhd->ibhd", pr, vf)
    o_ref.backward(dout.view(T, B, H, Dh).float())

    def close(a, b, name, tol=0.04):
        err = (a.float() - b).abs().max().item()
        sc = b.abs().max().item() + 1e-6
        assert err <= tol * sc + 2e-3, (name, err, sc)

    close(dq.view(T, B, H, Dh), quf.grad + qvf.grad, "dq")
    close(dkv[:, :, 0], kf.grad, "dk")
    close(dkv[:, :, 1], vf.grad, "dv")
    close(dr.view(K, H, Dh), rf.grad, "dr")
    close(du, quf.grad.sum((0, 1)), "du")
    close(dvb, qvf.grad.sum((0, 1)), "dvb")


# ---- the shapes bench.py times (BASELINE configs[1]: T = M = 2048, a 16 x 32 tile walk with distance blocks up to
# delta = 4095) and configs[4] (T = M = 4096, K = 8192), incl. memory reset and the same_length window ----
BIG_CASES = [
    (2048, 2048, 1, 2, 0, 2048, 0),
    (2048, 2048, 2, 1, 0, 2048, 1),
    (2048, 2048, 1, 1, 1, 2048, 0),
    (4096, 4096, 1, 1, 0, 4096, 0),
    (4096, 4096, 1, 1, 1, 4096, 1),
]


@pytest.mark.parametrize("T,M,B,H,same_length,mem_len,with_reset", BIG_CASES)
def test_relattn_fwd_bench_shapes(T, M, B, H, same_length, mem_len, with_reset):
    test_relattn_fwd(T, M, B, H, same_length, mem_len, with_reset, "commu_relattn_fwd_tc")


@pytest.mark.parametrize("mat", [True, False])
@pytest.mark.parametrize("T,M,B,H,same_length,mem_len,with_reset", BIG_CASES)
def test_relattn_bwd_bench_shapes(T, M, B, H, same_length, mem_len, with_reset, mat):
    from commu import _native as nv
    _relattn_bwd_case(nv, T, M, B, H, same_length, mem_len, with_reset, fwd_impl="commu_relattn_fwd_tc", mat=mat)


@pytest.mark.parametrize("r_std", [4.0, 8.0])
def test_relattn_fwd_large_position_scores(r_std):
    """|BD| of 100-300 (a trained checkpoint with strong position scores): the kernels stage the sheared position
    term as fp16 (ulp 0.125 at 128-256, i.e. <= 0.011 in the exponent after the 1/sqrt(64) scale and log2 e), so the
    probabilities stay within ~1 % and nothing overflows (fp16 max 65504).  Tolerances scaled accordingly."""
    from commu import _native as nv
    torch.manual_seed(5)
    dev = "cuda"
    T, M, B, H, Dh = 256, 384, 2, 2, 64
    K = T + M
    q = torch.randn(T, B, H, Dh, device=dev).bfloat16()
    kv = (torch.randn(K, B, 2, H, Dh, device=dev) * 0.5).bfloat16()
    r = (torch.randn(K, H, Dh, device=dev) * r_std).bfloat16()
    u = torch.randn(H, Dh, device=dev) * 0.5
    vb = torch.randn(H, Dh, device=dev) * 0.5
    scale = 1.0 / math.sqrt(Dh)
    out = torch.zeros(T, B, H * Dh, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    k_t, v_t = kv[:, :, 0], kv[:, :, 1]
    nv.call("commu_relattn_fwd_tc", q, H * Dh, k_t, v_t, 2 * H * Dh, r, H * Dh, K, u, vb, None,
            T, M, B, H, 0, T, scale, out, H * Dh, lse, None, None, None, None)
    torch.cuda.synchronize()
    ref, ref_lse, _, s = _relattn_ref(q.float(), k_t.float(), v_t.float(), r.float(), u, vb, None, T, M,
                                      B, H, Dh, 0, T, scale)
    bd_max = float((s[torch.isfinite(s)].abs().max()) / scale)
    assert bd_max > 12 * r_std, bd_max           # the case really has large raw scores
    assert torch.isfinite(out.float()).all()
    err = (out.view(T, B, H, Dh).float() - ref).abs().max().item()
    lerr = (lse - ref_lse).abs().max().item()
    assert err < 0.05, (err, bd_max)             # bf16 P / output rounding + fp16 position staging
    assert lerr < 0.03, (lerr, bd_max)


def test_embed_pos_ln_nll():
    from commu import _native as nv
    torch.manual_seed(3)
    dev = "cuda"
    d, dp, V, n = 60, 64, 53, 300
    table = torch.randn(V, d, device=dev)
    tok = torch.randint(0, V, (n,), device=dev)
    xf = torch.full((n, dp), 7.0, device=dev)
    xb = torch.full((n, dp), 7.0, device=dev, dtype=torch.bfloat16)
    nv.call("commu_embed_fwd", tok, table, d, dp, math.sqrt(d), n, xf, dp, xb, dp)
    ref = table[tok] * math.sqrt(d)
    assert torch.equal(xf[:, :d], ref) and xf[:, d:].abs().max() == 0
    assert torch.equal(xb[:, :d], ref.bfloat16())
    # embed bwd
    dx = torch.randn(n, dp, device=dev)
    dt = torch.zeros(V, d, device=dev)
    nv.call("commu_embed_bwd", tok, dx, dp, d, math.sqrt(d), n, dt)
    rdt = torch.zeros(V, d, device=dev).index_add_(0, tok, dx[:, :d] * math.sqrt(d))
    assert (dt - rdt).abs().max() < 1e-3
    # pos table
    inv = 1.0 / (10000 ** (torch.arange(0.0, d, 2.0, device=dev) / d))
    K = 77
    pf = torch.empty(K, dp, device=dev)
    pb = torch.empty(K, dp, device=dev, dtype=torch.bfloat16)
    nv.call("commu_pos_table", inv, K, 20, d, dp, pb, pf)
    dist = torch.arange(K, device=dev, dtype=torch.float32).clamp(max=20)
    ang = torch.outer(dist, inv)
    rp = torch.cat([ang.sin(), ang.cos()], -1)
    assert (pf[:, :d] - rp).abs().max() < 2e-6 and pf[:, d:].abs().max() == 0
    # layernorm fwd/bwd
    z = torch.randn(n, dp, device=dev)
    z[:, d:] = 0
    gam = torch.randn(d, device=dev)
    bet = torch.randn(d, device=dev)
    yf = torch.empty(n, dp, device=dev)
    yb = torch.empty(n, dp, device=dev, dtype=torch.bfloat16)
    mean = torch.empty(n, device=dev)
    rstd = torch.empty(n, device=dev)
    nv.call("commu_layernorm_fwd", z, dp, gam, bet, d, dp, 1e-5, n, yf, dp, yb, dp, mean, rstd)
    zz = z[:, :d].clone().requires_grad_(True)
    g2 = gam.clone().requires_grad_(True)
    b2 = bet.clone().requires_grad_(True)
    yr = torch.nn.functional.layer_norm(zz, (d,), g2, b2, 1e-5)
    assert (yf[:, :d] - yr).abs().max() < 1e-5
    dy = torch.randn(n, dp, device=dev)
    yr.backward(dy[:, :d])
    dzf = torch.empty(n, dp, device=dev)
    dzb = torch.empty(n, dp, device=dev, dtype=torch.bfloat16)
    dg = torch.zeros(d, device=dev)
    db = torch.zeros(d, device=dev)
    nv.call("commu_layernorm_bwd", dy, dp, z, dp, mean, rstd, gam, d, dp, n, dzf, dp, dzb, dp, dg, db, 0.0, 0)
    assert (dzf[:, :d] - zz.grad).abs().max() < 1e-4
    assert (dg - g2.grad).abs().max() < 1e-3 and (db - b2.grad).abs().max() < 1e-3
    assert (dzb[:, :d].float() - zz.grad).abs().max() < 0.02 * zz.grad.abs().max() + 1e-3
    # fused dropout of the bf16 gradient copy (the fp32 copy stays undropped)
    from helpers import drop_keep_mask, drop_keep_prob
    dzf2 = torch.empty(n, dp, device=dev)
    dzb2 = torch.empty(n, dp, device=dev, dtype=torch.bfloat16)
    dg2 = torch.zeros(d, device=dev)
    db2 = torch.zeros(d, device=dev)
    nv.call("commu_layernorm_bwd", dy, dp, z, dp, mean, rstd, gam, d, dp, n, dzf2, dp, dzb2, dp, dg2, db2, 0.25, 987654321)
    keep = drop_keep_mask(987654321, n, dp, 0.25).to(dev)[:, :d]
    assert torch.equal(dzf2, dzf)
    refd = zz.grad * keep / drop_keep_prob(0.25)
    assert (dzb2[:, :d].float() - refd).abs().max() < 0.02 * refd.abs().max() + 1e-3
    # nll
    Vp = 64
    lg = torch.randn(n, Vp, device=dev) * 3
    tgt = torch.randint(0, V, (n,), device=dev)
    nll = torch.empty(n, device=dev)
    lse = torch.empty(n, device=dev)
    nv.call("commu_nll_fwd", lg, Vp, V, tgt, n, nll, lse)
    lref = lg[:, :V].clone().requires_grad_(True)
    nref = torch.nn.functional.cross_entropy(lref, tgt, reduction="none")
    assert (nll - nref).abs().max() < 1e-5
    dl = torch.randn(n, device=dev)
    nref.backward(dl)
    dlg = torch.empty(n, Vp, device=dev, dtype=torch.bfloat16)
    nv.call("commu_nll_bwd", lg, Vp, V, Vp, lse, tgt, dl, n, dlg, Vp)
    assert (dlg[:, :V].float() - lref.grad).abs().max() < 0.02 and dlg[:, V:].abs().max() == 0


def test_cast_colsum_adam():
    from commu import _native as nv
    torch.manual_seed(4)
    dev = "cuda"
    # cast with head padding on rows (Dh=10 -> 64) and transpose
    H, Dh, d = 6, 10, 60
    w = torch.randn(H * Dh, d, device=dev)
    dst = torch.zeros(H * 64, 64, device=dev, dtype=torch.bfloat16)
    nv.call("commu_cast_pad", w, d, H * Dh, d, Dh, 64, d, 64, dst, 64, 0)
    ref = torch.zeros(H, 64, 64, device=dev)
    ref[:, :Dh, :d] = w.view(H, Dh, d)
    assert torch.equal(dst.view(H, 64, 64), ref.bfloat16())
    dstT = torch.zeros(64, H * 64, device=dev, dtype=torch.bfloat16)
    nv.call("commu_cast_pad", w, d, H * Dh, d, Dh, 64, d, 64, dstT, H * 64, 1)
    assert torch.equal(dstT, ref.view(H * 64, 64).t().contiguous().bfloat16())
    # unpad accumulate
    gp = torch.randn(H * 64, 64, device=dev)
    acc = torch.ones(H * Dh, d, device=dev)
    nv.call("commu_unpad_accum", gp, 64, H * Dh, d, Dh, 64, d, 64, acc, d, 0.5)
    assert (acc - (1 + 0.5 * gp.view(H, 64, 64)[:, :Dh, :d].reshape(H * Dh, d))).abs().max() < 1e-6
    # colsum
    x = torch.randn(1000, 136, device=dev).bfloat16()
    out = torch.zeros(130, device=dev)
    nv.call("commu_colsum_bf16", x, 136, 130, 1000, out)
    assert (out - x.float()[:, :130].sum(0)).abs().max() < 1e-2
    x8 = torch.randn(1003, 520, device=dev).bfloat16()          # vector path: 8 columns per thread, row tail
    out8 = torch.ones(512, device=dev)
    nv.call("commu_colsum_bf16", x8, 520, 512, 1003, out8)
    assert (out8 - 1 - x8.float()[:, :512].sum(0)).abs().max() < 2e-2
    # clip + adam vs torch
    n = 100003
    n_al = (n + 3) // 4 * 4
    p = torch.randn(n_al, device=dev)
    p0 = p.clone()
    gbuf = torch.randn(n_al, device=dev) * 0.1
    gbuf[n:] = 0
    m = torch.zeros(n_al, device=dev)
    v = torch.zeros(n_al, device=dev)
    ref_p = torch.nn.Parameter(p0[:n].clone())
    wd = 0.01                                     # Adam's L2 term (cfg.TRAIN.weight_decay, train.py:442-443)
    opt = torch.optim.Adam([ref_p], lr=0.004, weight_decay=wd)
    gn = torch.zeros(1, device=dev)
    gout = torch.zeros(1, device=dev)
    for step in range(1, 4):
        g = gbuf[:n] * step
        ref_p.grad = g.clone()
        tn = torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        gb = gbuf * step
        gn.zero_()
        nv.call("commu_sumsq", gb, n_al, gn)
        pb = torch.empty(n_al, device=dev, dtype=torch.bfloat16)
        nv.call("commu_clip_adam", p, gb, m, v, n_al, 0.004, 0.9, 0.999, 1e-8, step, gn, 1.0, 1.0, wd, gout, pb)
        assert torch.equal(pb, p.bfloat16())
        assert abs(gout.item() - tn.item()) / tn.item() < 1e-4
        assert (p[:n] - ref_p.data).abs().max() < 2e-6
