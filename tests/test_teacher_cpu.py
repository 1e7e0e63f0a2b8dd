"""Host-side generation loop + chord teacher forcing against sequences produced by the
reference's own InferenceTask.generate_sequence / TeacherForceTask driven by the same deterministic
stand-in for the network (tests/golden/make_golden.py::fake_model_patch).  CPU only: the sampler
arithmetic is stubbed with torch ops here (the CUDA sampler has its own GPU parity tests)."""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

from helpers import GOLDEN

sys.path.insert(0, GOLDEN)
from make_golden import TEACHER_SCENARIOS, fake_model_patch  # noqa: E402


def _cpu_sampler(task):
    def calc_probs(logits):
        if task.input_data.temperature == 0:
            probs = torch.zeros_like(logits)
            probs[logits.argmax()] = 1.0
        else:
            logits /= task.input_data.temperature          # in place (quirk Q3)
            probs = F.softmax(logits, dim=-1)
        return F.pad(probs, [1, 0])

    def apply_sampling(probs, wrong_tokens):
        _, idx = torch.topk(probs, task.input_data.top_k)
        mask = torch.zeros_like(probs)
        mask[idx] = 1.0
        for w in wrong_tokens or []:
            mask[w] = 0.0
        probs *= mask
        probs /= probs.sum()
        return probs

    task.calc_probs = calc_probs
    task.apply_sampling = apply_sampling


def test_generation_loop_matches_reference():
    from commu.midi_generator.midi_inferrer import InferenceTask, TeacherForceTask
    z = np.load(os.path.join(GOLDEN, "teacher_forcing.npz"))
    for name, (nm, ctok, cpos, script) in TEACHER_SCENARIOS.items():
        task = InferenceTask(torch.device("cpu"))
        task.input_data = SimpleNamespace(num_measures=nm, temperature=0.95, top_k=32, num_generate=1,
                                          chord_token_components={"chord_token": list(ctok), "chord_position": list(cpos)})
        task.inference_cfg = SimpleNamespace(GENERATION=SimpleNamespace(generation_length=200))
        fake_model_patch(task, script)
        _cpu_sampler(task)
        seq = task.generate_sequence([0, 574, 623, 627, 635, 639, 642, 651, 684, 694, 720, 727], 0)
        assert seq is not None, name
        assert seq == z[name].tolist(), name
        assert task.validate_generated_sequence(seq)
        t = TeacherForceTask(task.input_data)
        assert t.chord_token == list(ctok) and t.chord_position == list(cpos)


def test_teacher_predicates():
    from commu.midi_generator.midi_inferrer import TeacherForceTask
    d = SimpleNamespace(num_measures=4, chord_token_components={"chord_token": [200, 201], "chord_position": [432, 496]})
    t = TeacherForceTask(d)
    assert t.inter_chord_flags == [False, True] and not t.check_length_fit()
    assert t.check_first_position([0, 2]) and not t.check_first_position([0, 3])
    assert t.check_mul_chord_per_bar_case([0, 2, 432])
    t.teach_chord_token()
    assert t.next_tokens_forced == [200] and t.chord_token == [201]
    assert t.check_chord_position_passed(500) and t.check_chord_position_passed(2) and not t.check_chord_position_passed(450)
    assert t.check_wrong_chord_token_generated(195) and t.check_wrong_eos_generated(1)
    t.teach_remnant_chord()
    assert t.next_tokens_forced[-1] == 496


def _scripted_sampler(script_state, script, logits, temperature, k, top_k, wrong):
    """What the sampler kernel does with one row (softmax(l / T^k) -> top-k -> wrong-token mask -> renormalise),
    followed by the scripted 'draw' of the golden generator (arg-max; advances the script pointer)."""
    lg = logits / (temperature ** k)
    probs = F.pad(F.softmax(lg, dim=-1), [1, 0])
    _, idx = torch.topk(probs, top_k)
    mask = torch.zeros_like(probs)
    mask[idx] = 1.0
    for w in wrong or []:
        mask[w] = 0.0
    probs = probs * mask
    probs = probs / probs.sum()
    script_state["ptr"] += 1
    return int(torch.argmax(probs))


def test_sequence_coroutine_matches_reference():
    """The per-sequence coroutine of the batched generation (InferenceTask._sequence_steps) asks for model steps and
    samples at exactly the points the reference loop does: driven by the golden generator's scripted stand-in it
    reproduces the reference sequences (teacher forcing, rejected chord tokens with the logits re-used and divided
    by the temperature again, forced tokens fed twice)."""
    from commu.midi_generator.midi_inferrer import InferenceTask
    z = np.load(os.path.join(GOLDEN, "teacher_forcing.npz"))
    for name, (nm, ctok, cpos, script) in TEACHER_SCENARIOS.items():
        task = InferenceTask(torch.device("cpu"))
        task.input_data = SimpleNamespace(num_measures=nm, temperature=0.95, top_k=32, num_generate=1,
                                          chord_token_components={"chord_token": list(ctok), "chord_position": list(cpos)})
        task.inference_cfg = SimpleNamespace(GENERATION=SimpleNamespace(generation_length=200))
        st = fake_model_patch(task, script)
        gen = task._sequence_steps([0, 574, 623, 627, 635, 639, 642, 651, 684, 694, 720, 727])
        logits, seq, n_first = None, None, 0
        try:
            req = next(gen)
            while True:
                if req[0] == "step":
                    n_first += int(req[2] is False)
                    logits, _ = task.calc_logits_and_mems(None, 0)
                    req = gen.send(None)
                else:
                    tok = _scripted_sampler(st, script, logits, 0.95, req[1], 32, req[2])
                    req = gen.send(tok)
        except StopIteration as fin:
            seq = fin.value
        assert n_first == 1, name                     # Q1: exactly one step whose memory is dropped
        assert seq == z[name].tolist(), name


def test_generate_batch_lockstep(monkeypatch):
    """generate_batch: three sequences with different scripts share one (stubbed) batch engine; every live sequence
    gets exactly one token step per engine step, the first step's state is dropped for the whole batch, rejected
    tokens are re-drawn in extra sampler waves with the row's wrong-token mask and temperature ** k."""
    import commu.midi_generator.midi_inferrer as mi
    z = np.load(os.path.join(GOLDEN, "teacher_forcing.npz"))
    names = list(TEACHER_SCENARIOS)
    nm, ctok, cpos, script = TEACHER_SCENARIOS[names[0]]
    B, V = 3, 729
    # all rows run scenario 0's metadata (one input_data per call, like the reference) but row-specific noise
    scripts = [script, script, script]
    states = [{"calls": 0, "ptr": 0} for _ in range(B)]

    class FakeEngine:
        def __init__(self):
            self.B, self.V, self.steps, self.rewinds = B, V, 0, 0
            self.logits = torch.zeros(B, V)

        def step(self, tokens, state):
            assert tokens.shape == (B,)
            self.steps += 1
            if state.count < 11:                      # memory pre-fill of [0] + meta[:-1]: not a scripted call
                return self.logits, mi.DecodeState(state.count + 1, state.slot + 1)
            for r in range(B):
                states[r]["calls"] += 1
                g = torch.Generator().manual_seed(1000 + states[r]["calls"])
                base = torch.randn(728, generator=g)
                want = scripts[r][states[r]["ptr"]] if states[r]["ptr"] < len(scripts[r]) else 1
                base[want - 1] += 20.0
                self.logits[r, 0] = 0.0
                self.logits[r, 1:] = base
            return self.logits, mi.DecodeState(state.count + 1, state.slot + 1)

    def fake_call(name, logits, ld, b, v, temp, top_k, top_p, wrong, seed, offset, toks, probs, ldp, dstate):
        assert name == "commu_sample" and b == B
        for r in range(B):
            w = torch.nonzero(wrong[r]).flatten().tolist()
            lg = logits[r, 1:] / temp
            pr = F.pad(F.softmax(lg, dim=-1), [1, 0])
            _, idx = torch.topk(pr, top_k)
            mask = torch.zeros_like(pr)
            mask[idx] = 1.0
            mask[w] = 0.0
            toks[r] = int(torch.argmax(pr * mask))
    task = mi.InferenceTask(torch.device("cpu"))
    task.input_data = SimpleNamespace(num_measures=nm, temperature=0.95, top_k=32, num_generate=B,
                                      chord_token_components={"chord_token": list(ctok), "chord_position": list(cpos)})
    task.inference_cfg = SimpleNamespace(GENERATION=SimpleNamespace(generation_length=200))
    eng = FakeEngine()
    monkeypatch.setattr(task, "_batch_engine", lambda n: eng)
    monkeypatch.setattr(mi.nv, "call", fake_call)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    orig_adv = None
    # the scripted draw advances the row's pointer whenever the row consumed a sample: emulate through a wrapper
    real = task._sequence_steps

    def counting(seq):
        row = counting.n
        counting.n += 1
        gen = real(seq)
        try:
            req = next(gen)
            while True:
                val = yield req
                if req[0] == "sample":
                    states[row]["ptr"] += 1
                req = gen.send(val)
        except StopIteration as fin:
            return fin.value
    counting.n = 0
    monkeypatch.setattr(task, "_sequence_steps", counting)
    meta = [574, 623, 627, 635, 639, 642, 651, 684, 694, 720, 727]
    out = task.generate_batch(meta, B)
    assert all(s is not None for s in out)
    for s in out:
        assert s == z[names[0]].tolist()
    assert task.validate_generated_sequence(out[0])
