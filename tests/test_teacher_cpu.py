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
