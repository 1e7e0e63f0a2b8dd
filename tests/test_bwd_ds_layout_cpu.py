"""Index conventions of the materialised-dS attention backward planned in DESIGN.md section 7 (1a), pinned on the CPU
against autograd of the oracle's attention formulas (oracle/transfoxl_oracle.py `_layer`):

  dS[b,h,i,j]                      gradient of the pre-softmax score (before the 1/sqrt(Dh) scale is applied to the
                                   operands), stored key-indexed;
  dB[b,h,i, j - i + T - 1] = dS    the same values with the relative shift done by the ADDRESS (row i shifted by
                                   T-1-i columns): its columns are a reversed distance axis, dp = M+T-1 - delta; the
                                   causal limit j <= i + M keeps dp inside [0, K-1], so dB is T x K like dS (the
                                   masked upper triangle, which would fall beyond column K-1, is never stored);
  dq   = scale * ( dS K  +  dB R' ),           R'[dp] = R[M+T-1 - dp]      (two plain products, no shear)
  dR'  = scale * sum_{b,i} dB[b,h,i,:]^T (q_i + v),   dR[delta] = dR'[M+T-1 - delta]
  d r_r_bias = scale * sum dB R' ;  d r_w_bias = scale * sum dS K   (row sums of the same products)
"""
import math

import torch

from helpers import orc  # noqa: F401  (puts the repo root on sys.path like the other tests)


def test_skewed_ds_layout_reproduces_autograd():
    torch.manual_seed(0)
    T, M, B, H, Dh = 7, 5, 2, 3, 4
    K = T + M
    scale = 1.0 / math.sqrt(Dh)
    dd = torch.float64
    q = torch.randn(T, B, H, Dh, dtype=dd, requires_grad=True)
    k = torch.randn(K, B, H, Dh, dtype=dd, requires_grad=True)
    v = torch.randn(K, B, H, Dh, dtype=dd, requires_grad=True)
    R = torch.randn(K, H, Dh, dtype=dd, requires_grad=True)        # by distance, as in the oracle
    u = torch.randn(H, Dh, dtype=dd, requires_grad=True)
    vb = torch.randn(H, Dh, dtype=dd, requires_grad=True)
    ii = torch.arange(T)[:, None]
    jj = torch.arange(K)[None, :]
    valid = (jj <= ii + M)
    AC = torch.einsum("ibhd,jbhd->bhij", q + u, k)
    QR = torch.einsum("ibhd,thd->bhit", q + vb, R)
    dist = (ii + M - jj).clamp(min=0)
    BD = QR.gather(3, dist[None, None].expand(B, H, T, K))
    raw = AC + BD
    raw.retain_grad()
    score = (raw * scale).masked_fill(~valid[None, None], float("-inf"))
    prob = torch.softmax(score, -1)
    av = torch.einsum("bhij,jbhd->ibhd", prob, v)
    (av * torch.randn_like(av)).sum().backward()
    dS = raw.grad                                                  # [B,H,T,K]; already includes the scale factor
    assert float(dS[:, :, ~valid].abs().max()) == 0.0             # masked scores get no gradient

    # address-sheared copy: row i shifted right by T-1-i columns; only visible keys (j <= i + M) are stored, which
    # keeps every column index below K; everything never written stays zero
    dB = torch.zeros(B, H, T, K, dtype=dd)
    for i in range(T):
        nvis = i + M + 1                                           # keys 0 .. i+M
        dB[:, :, i, T - 1 - i: T - 1 - i + nvis] = dS[:, :, i, :nvis]
        assert T - 1 - i + nvis == K
    # dp = j - i + T - 1 = (M + T - 1) - delta with delta = i + M - j  ->  R'[dp] = R[K-1 - dp]: R reversed
    Rp = R.detach().flip(0)
    qv = (q + vb).detach()
    qu = (q + u).detach()
    # dq
    dq_key = torch.einsum("bhij,jbhd->ibhd", dS, k.detach())
    dq_pos = torch.einsum("bhip,phd->ibhd", dB, Rp)
    assert torch.allclose(dq_key + dq_pos, q.grad, atol=1e-10)
    # dR through the reversed axis
    dRp = torch.einsum("bhip,ibhd->phd", dB, qv)
    assert torch.allclose(dRp.flip(0), R.grad, atol=1e-10)
    # biases
    assert torch.allclose(dq_key.sum((0, 1)), u.grad, atol=1e-10)
    assert torch.allclose(dq_pos.sum((0, 1)), vb.grad, atol=1e-10)
    # dk from the key-indexed copy (what the dk/dv pass already does)
    assert torch.allclose(torch.einsum("bhij,ibhd->jbhd", dS, qu), k.grad, atol=1e-10)
