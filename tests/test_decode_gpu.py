"""GPU parity of the decode engine and the sampler against the reference-generated goldens and the
oracle."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from helpers import GOLDEN, load_golden, orc
from test_model_gpu import build_model

pytestmark = pytest.mark.gpu


def _run_decode(precision):
    from commu.engine.decode import DecodeEngine
    z, cfg, P = load_golden("decode_greedy")
    model = build_model(cfg, P)
    model.eval()
    eng = DecodeEngine(model, batch=2, mem_len=cfg.mem_len, same_length=True, precision=precision)
    ctx = torch.from_numpy(z["ctx"]).cuda()
    _, st = eng.prefill(ctx[:-1])
    cur = ctx[-1].contiguous()
    toks, worst = [], 0.0
    for t in range(z["tokens"].shape[0]):
        lg, st = eng.step(cur, st)
        ref = torch.from_numpy(z["logits"][t])
        worst = max(worst, float((lg.cpu() - ref).abs().max()))
        nxt, _ = eng.sample(lg, 0.0)                      # greedy on device
        toks.append(nxt.cpu().numpy())
        cur = torch.from_numpy(z["tokens"][t]).cuda()     # follow the reference trajectory
    return np.stack(toks), worst, z


def test_decode_fp32_greedy_tokens_identical():
    toks, worst, z = _run_decode("fp32")
    assert worst < 2e-4, worst
    assert np.array_equal(toks, z["tokens"])


def test_decode_fp32_free_running_greedy_identical():
    """No teacher forcing: the engine's own greedy tokens are fed back (>= mem_len steps so the ring wraps)."""
    from commu.engine.decode import DecodeEngine
    z, cfg, P = load_golden("decode_greedy")
    model = build_model(cfg, P)
    eng = DecodeEngine(model, batch=2, mem_len=cfg.mem_len, same_length=True, precision="fp32")
    ctx = torch.from_numpy(z["ctx"]).cuda()
    _, st = eng.prefill(ctx[:-1])
    cur = ctx[-1].contiguous()
    for t in range(z["tokens"].shape[0]):
        lg, st = eng.step(cur, st)
        cur, _ = eng.sample(lg, 0.0)
        assert np.array_equal(cur.cpu().numpy(), z["tokens"][t]), t


def test_decode_bf16_logits_close():
    toks, worst, z = _run_decode("bf16")
    assert worst < 0.06 * np.abs(z["logits"]).max() + 0.02, worst
    # tokens must agree wherever the reference's greedy margin is not razor thin
    srt = np.sort(z["logits"][:, :, 1:], axis=-1)
    margin = srt[..., -1] - srt[..., -2]
    ok = margin > 4 * worst
    assert np.array_equal(toks[ok], z["tokens"][ok])


def test_sampler_matches_reference_probs():
    from commu import _native as nv
    z = np.load(os.path.join(GOLDEN, "sampler_probs.npz"))
    for ci in range(4):
        full = torch.from_numpy(z["case%d/logits_full" % ci]).cuda().unsqueeze(0).contiguous()
        temp, top_k = z["case%d/params" % ci]
        wrong = z["case%d/wrong" % ci]
        V = full.shape[1]
        wr = None
        if len(wrong):
            wr = torch.zeros(1, V, dtype=torch.uint8, device="cuda")
            wr[0, torch.from_numpy(wrong).cuda()] = 1
        probs = torch.empty(1, V, device="cuda")
        nv.call("commu_sample", full, V, 1, V, float(temp), int(top_k), 0.0, wr, 0, 0, None, probs, V, None)
        ref = z["case%d/probs" % ci]
        got = probs[0].cpu().numpy()
        assert np.array_equal(got > 0, ref > 0), ci
        assert np.abs(got - ref).max() < 1e-6, ci


def test_sampler_top_p_and_draws():
    from commu import _native as nv
    torch.manual_seed(0)
    V, B = 729, 64
    lg = (torch.randn(B, V) * 2).cuda()
    probs = torch.empty(B, V, device="cuda")
    toks = torch.empty(B, dtype=torch.int64, device="cuda")
    nv.call("commu_sample", lg, V, B, V, 0.95, 0, 0.9, None, 123, 7, toks, probs, V, None)
    for b in range(0, B, 9):
        ref = orc.sampler_probs(lg[b].cpu(), 0.95, 0, 0.9, [])
        got = probs[b].cpu()
        assert torch.equal(got > 0, ref > 0), b
        assert (got - ref).abs().max() < 1e-5
        assert got[toks[b]] > 0 and toks[b] >= 1
    # empirical distribution of the counter-based draw follows the probabilities
    row = lg[:1].contiguous()
    p1 = torch.empty(1, V, device="cuda")
    counts = torch.zeros(V)
    t1 = torch.empty(1, dtype=torch.int64, device="cuda")
    n = 4000
    for i in range(n):
        nv.call("commu_sample", row, V, 1, V, 1.0, 8, 0.0, None, 99, i, t1, p1, V, None)
        counts[int(t1)] += 1
    pr = p1[0].cpu()
    assert set(torch.nonzero(counts).flatten().tolist()) <= set(torch.nonzero(pr).flatten().tolist())
    assert (counts / n - pr).abs().max() < 0.03


def test_generate_graph_equals_eager_greedy():
    """CUDA-graph replay of the decode step (device-resident ring state) == eager per-step launches."""
    from commu.engine.decode import DecodeEngine
    z, cfg, P = load_golden("decode_greedy")
    model = build_model(cfg, P)
    ctx = torch.from_numpy(z["ctx"]).cuda()
    eng = DecodeEngine(model, batch=2, mem_len=cfg.mem_len, same_length=True, precision="fp32")
    a = eng.generate(ctx, 40, temperature=0.0, top_k=0, top_p=0.0, use_graph=False).cpu().numpy()
    eng2 = DecodeEngine(model, batch=2, mem_len=cfg.mem_len, same_length=True, precision="fp32")
    b = eng2.generate(ctx, 40, temperature=0.0, top_k=0, top_p=0.0, use_graph=True).cpu().numpy()
    assert np.array_equal(a, z["tokens"])
    assert np.array_equal(b, z["tokens"])


def test_inference_task_surface():
    from commu.midi_generator.midi_inferrer import InferenceTask
    z, cfg, P = load_golden("decode_greedy")
    model = build_model(cfg, P)
    model.eval()
    model.reset_length(1, cfg.mem_len)
    task = InferenceTask(torch.device("cuda"))
    task(model, NS(temperature=0.0, top_k=32), None)
    meta = [int(t) for t in z["ctx"][1:, 0]]
    seq, mems = task.init_seq_and_mems(meta, len(meta))
    assert seq == [0] + meta
    logits, mems2 = task.calc_logits_and_mems(seq, mems)
    assert logits.shape[0] == cfg.n_token - 1
    probs = task.apply_sampling(task.calc_probs(logits), [])
    tok = task.infer_token(probs)
    assert tok == 1 + int(logits.argmax())
