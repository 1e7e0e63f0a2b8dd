"""Pins the oracle (oracle/transfoxl_oracle.py) against outputs of the reference itself, frozen in
tests/golden/*.npz by tests/golden/make_golden.py (CPU only)."""
import numpy as np
import torch

from helpers import load_golden, orc, rel_err


def _run_segments(name, dtype):
    z, cfg, P = load_golden(name)
    P = {k: v.to(dtype).requires_grad_(True) for k, v in P.items()}
    mems = None
    nseg = len([k for k in z.files if k.endswith("/loss")])
    for s in range(nseg):
        data = torch.from_numpy(z["seg%d/data" % s])
        target = torch.from_numpy(z["seg%d/target" % s])
        reset = torch.from_numpy(z["seg%d/reset" % s])
        loss, mems = orc.forward_loss(cfg, P, data, target, reset, mems)
        loss.mean().backward()
        assert rel_err(loss.detach(), z["seg%d/loss" % s]) < 2e-5, (name, s)
        assert mems.shape == z["seg%d/mems" % s].shape
        assert rel_err(mems, z["seg%d/mems" % s]) < 2e-5, (name, s)
    for k, p in P.items():
        g = z["grad/" + k]
        assert rel_err(p.grad, g) < 5e-4, (name, k, rel_err(p.grad, g))


def test_forward_backward_basic_fp32():
    _run_segments("fwd_basic", torch.float32)


def test_forward_backward_samelen_clamp_dh10_fp32():
    _run_segments("fwd_samelen_dh10", torch.float32)


def test_forward_backward_fp64_close_to_reference_fp32():
    _run_segments("fwd_basic", torch.float64)


def test_decode_greedy_tokens_identical():
    z, cfg, P = load_golden("decode_greedy")
    ctx = torch.from_numpy(z["ctx"])
    with torch.no_grad():
        _, mems = orc.forward_generate(cfg, P, ctx[:-1], None)
        cur = ctx[-1:]
        for t in range(z["tokens"].shape[0]):
            lg, mems = orc.forward_generate(cfg, P, cur, mems)
            assert rel_err(lg[-1], z["logits"][t]) < 2e-5
            nxt = torch.tensor([orc.greedy_token(lg[-1, b]) for b in range(lg.shape[1])])
            assert np.array_equal(nxt.numpy(), z["tokens"][t]), t
            cur = nxt[None]
    assert rel_err(mems, z["final_mems"]) < 2e-5


def test_sampler_probs():
    z = np.load(__import__("os").path.join(__import__("helpers").GOLDEN, "sampler_probs.npz"))
    for ci in range(4):
        full = torch.from_numpy(z["case%d/logits_full" % ci])
        temp, top_k = z["case%d/params" % ci]
        wrong = list(z["case%d/wrong" % ci])
        p = orc.sampler_probs(full, float(temp), int(top_k), 0.0, wrong)
        ref = z["case%d/probs" % ci]
        assert np.array_equal(p.numpy() > 0, ref > 0)
        assert np.abs(p.numpy() - ref).max() < 1e-6


def test_top_p_definition():
    lg = torch.log(torch.tensor([1e-9, 0.5, 0.3, 0.15, 0.05]))
    p = orc.sampler_probs(lg, 1.0, 0, 0.9, [])
    assert p[0] == 0 and p[4] == 0 and abs(float(p[1:4].sum()) - 1) < 1e-6
    assert abs(float(p[1]) - 0.5 / 0.95) < 1e-5


def test_train_steps_match_reference_loop():
    z, cfg, P = load_golden("train_steps")
    lr, warmup, lr_min, clip, chunks = z["hyper"]
    chunks = int(chunks)
    opt = orc.AdamState(P)
    mems = [None] * chunks
    nsteps = len(z["losses"])
    for s in range(nsteps):
        data = torch.from_numpy(z["step%d/data" % s])
        target = torch.from_numpy(z["step%d/target" % s])
        reset = torch.from_numpy(z["step%d/reset" % s])
        batches = list(zip(torch.chunk(data, chunks, 1), torch.chunk(target, chunks, 1),
                           torch.chunk(reset, chunks, 0)))
        cur_lr = lr * orc.lr_multiplier(s, warmup, lr, lr_min)
        assert abs(cur_lr - z["lrs"][s]) < 1e-12
        loss, gn, mems, _ = orc.train_step(cfg, P, opt, batches, mems, cur_lr, clip=clip)
        assert abs(loss - z["losses"][s]) / z["losses"][s] < 1e-5, (s, loss, z["losses"][s])
        assert abs(gn - z["gnorms"][s]) / z["gnorms"][s] < 1e-3
    for k in P:
        assert rel_err(P[k], z["final/" + k]) < 2e-3, k
