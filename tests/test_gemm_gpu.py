"""GPU parity of the tcgen05 GEMM (commu_gemm_bf16) against torch fp32 matmul on the same
bf16-rounded operands, plus the naive SIMT kernel as an independent cross-check."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, b, a_mn, b_mn):
    A = a.float().t() if a_mn else a.float()
    B = b.float().t() if b_mn else b.float()
    return A @ B.t()


CASES = [
    # m, n, k, a_mn, b_mn, split
    (128, 128, 64, 0, 0, 1),
    (256, 128, 512, 0, 0, 1),
    (300, 200, 136, 0, 0, 1),      # ragged tiles, k tail
    (1024, 512, 512, 0, 0, 1),     # BLOCK_N = 256 path
    (777, 1536, 520, 0, 0, 1),
    (128, 128, 128, 1, 1, 1),      # wgrad majors
    (512, 512, 2048, 1, 1, 4),     # wgrad + split-k atomics
    (2048, 512, 4096, 1, 1, 8),
    (200, 328, 1000, 1, 1, 3),
    (256, 256, 256, 0, 1, 1),
    (256, 256, 256, 1, 0, 1),
]


@pytest.mark.parametrize("m,n,k,a_mn,b_mn,split", CASES)
@pytest.mark.parametrize("impl", [0, 1, 2, 3])
def test_gemm_plain(m, n, k, a_mn, b_mn, split, impl):
    from commu import _native as nv
    torch.manual_seed(m * 7 + n * 3 + k)
    dev = "cuda"
    ka = (k + 7) // 8 * 8
    ma = (m + 7) // 8 * 8
    na = (n + 7) // 8 * 8
    a = torch.randn((ka, ma) if a_mn else (m, ka), device=dev).bfloat16()
    b = torch.randn((ka, na) if b_mn else (n, ka), device=dev).bfloat16()
    a_l = a[:k, :m] if a_mn else a[:, :k]
    b_l = b[:k, :n] if b_mn else b[:, :k]
    ref = _ref(a_l, b_l, a_mn, b_mn)
    out = torch.zeros(m, na, device=dev, dtype=torch.float32)
    nv.gemm(a, b, m=m, n=n, k=k, a_mn=a_mn, b_mn=b_mn, split_k=split, out_f32=out,
            f32_atomic=split > 1, impl=impl)
    torch.cuda.synchronize()
    err = (out[:, :n] - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-3 * scale + 1e-3, (err, scale)
    if na > n:
        assert out[:, n:].abs().max().item() == 0.0


@pytest.mark.parametrize("impl", [0, 1, 2, 3])
def test_gemm_epilogue(impl):
    from commu import _native as nv
    torch.manual_seed(5)
    dev = "cuda"
    m, n, k = 384, 520, 256
    a = torch.randn(m, k, device=dev).bfloat16()
    b = torch.randn(n, k, device=dev).bfloat16()
    bias = torch.randn(n, device=dev)
    mask = torch.randn(m, n, device=dev).bfloat16()
    add = torch.randn(m, n, device=dev)
    acc = a.float() @ b.float().t()
    # bias + relu -> bf16
    o1 = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    nv.gemm(a, b, m=m, n=n, k=k, bias=bias, relu=True, out_bf16=o1, impl=impl)
    r1 = torch.relu(acc + bias)
    assert (o1.float() - r1).abs().max().item() <= 0.02 * r1.abs().max().item()
    # relu mask + add -> f32 and bf16 together, alpha
    o2 = torch.empty(m, n, device=dev)
    o2b = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    nv.gemm(a, b, m=m, n=n, k=k, alpha=0.5, relu_mask=mask, add_f32=add, out_f32=o2, out_bf16=o2b,
            impl=impl)
    r2 = (0.5 * acc) * (mask.float() > 0) + add
    assert (o2 - r2).abs().max().item() <= 2e-3 * r2.abs().max().item()
    assert (o2b.float() - r2).abs().max().item() <= 0.02 * r2.abs().max().item()
    # accumulate into existing f32 via atomic
    o3 = add.clone()
    nv.gemm(a, b, m=m, n=n, k=k, out_f32=o3, f32_atomic=True, impl=impl)
    assert (o3 - (acc + add)).abs().max().item() <= 2e-3 * acc.abs().max().item()


def test_gemm_perf(out_dir):
    """Not a pass/fail perf gate: records TFLOP/s of representative shapes for the round log."""
    import json, os
    from commu import _native as nv
    dev = "cuda"
    res = []
    shapes = [(32768, 512, 512, 0, 0, 1), (32768, 2048, 512, 0, 0, 1), (32768, 512, 2048, 0, 0, 1),
              (65536, 1024, 512, 0, 0, 1), (2048, 512, 32768, 1, 1, 16), (512, 512, 32768, 1, 1, 36),
              (8192, 8192, 8192, 0, 0, 1)]
    for (m, n, k, a_mn, b_mn, split) in shapes:
        a = torch.randn((k, m) if a_mn else (m, k), device=dev).bfloat16()
        b = torch.randn((k, n) if b_mn else (n, k), device=dev).bfloat16()
        if split > 1:
            out = torch.zeros(m, n, device=dev)
            kw = dict(out_f32=out, f32_atomic=True)
        else:
            out = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
            kw = dict(out_bf16=out)
        for _ in range(3):
            nv.gemm(a, b, m=m, n=n, k=k, a_mn=a_mn, b_mn=b_mn, split_k=split, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        iters = 10
        for _ in range(iters):
            nv.gemm(a, b, m=m, n=n, k=k, a_mn=a_mn, b_mn=b_mn, split_k=split, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        res.append(dict(m=m, n=n, k=k, a_mn=a_mn, b_mn=b_mn, split=split, ms=ms,
                        tflops=2.0 * m * n * k / ms / 1e9))
    with open(os.path.join(out_dir, "gemm_perf.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


@pytest.mark.parametrize("impl", [2, 3])
def test_gemm_dropout_epilogue(impl):
    """Fused inverted dropout of the GEMM result (bias -> relu -> dropout -> residual): the mask is the counter-based
    keep function of commu_dropout, restated in numpy in tests/helpers.py."""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import drop_keep_mask, drop_keep_prob
    from commu import _native as nv
    torch.manual_seed(3)
    dev = "cuda"
    m, n, k, p_drop, seed = 300, 320, 256, 0.1, 0x1234_5678_9ABC_DEF1
    a = torch.randn(m, k, device=dev).bfloat16()
    b = torch.randn(n, k, device=dev).bfloat16()
    bias = torch.randn(n, device=dev)
    add = torch.randn(m, n, device=dev)
    keep = drop_keep_mask(seed, m, n, p_drop).to(dev)
    ref = torch.relu(a.float() @ b.float().t() + bias) * keep / drop_keep_prob(p_drop) + add
    out = torch.empty(m, n, device=dev)
    outb = torch.empty(m, n, device=dev, dtype=torch.bfloat16)
    nv.gemm(a, b, m=m, n=n, k=k, bias=bias, relu=True, add_f32=add, out_f32=out, out_bf16=outb, impl=impl,
            drop_p=p_drop, drop_seed=seed)
    torch.cuda.synchronize()
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-3 * scale + 1e-3
    assert (outb.float() - ref).abs().max().item() <= 1e-2 * scale
    frac = 1.0 - keep.float().mean().item()
    assert abs(frac - p_drop) < 0.01
