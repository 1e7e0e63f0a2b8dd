"""GPU parity of the generation loop against the REFERENCE's own InferenceTask.generate_sequence run
on a real (tiny) reference model at temperature 0 (tests/golden/make_golden.py::gen_generate_case): chord teacher
forcing, forced first positions, inter-chord positions, end of sequence - every generated and forced token must be
identical, on the one-sequence path and on the batched path, through the fp32 decode engine (C-ABI)."""
import os
from types import SimpleNamespace as NS

import numpy as np
import pytest
import torch

from helpers import GOLDEN, orc
from test_model_gpu import build_model

pytestmark = pytest.mark.gpu

import sys
sys.path.insert(0, GOLDEN)
from make_golden import GENERATE_CFG, GENERATE_SCENARIOS  # noqa: E402


def _task(z, sc, n):
    from commu.midi_generator.midi_inferrer import InferenceTask
    seed, std, bar, nm, ctok, cpos, gen_len = GENERATE_SCENARIOS[sc]
    c = GENERATE_CFG
    cfg = orc.make_cfg(n_layer=c["n_layer"], n_head=c["n_head"], d_model=c["d_model"], d_inner=c["d_inner"], tgt_len=1,
                       mem_len=c["mem_len"], same_length=True, clamp_len=-1, n_token=729)
    P = {k[len(sc) + 7:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(sc + "/param/")}
    P.pop("pos_emb.inv_freq", None)
    model = build_model(cfg, P)
    model.eval()
    model.reset_length(1, c["mem_len"])
    task = InferenceTask(torch.device("cuda"))
    task(model, NS(num_measures=nm, temperature=0.0, top_k=32, num_generate=n,
                   chord_token_components={"chord_token": list(ctok), "chord_position": list(cpos)}),
         NS(GENERATION=NS(generation_length=gen_len)))
    return task


@pytest.fixture
def raw_sequences(monkeypatch):
    """Record the sequence before validate_teacher_forced_sequence judges it (as the golden generator does)."""
    import commu.midi_generator.midi_inferrer as mi
    seen = []
    orig = mi.TeacherForceTask.validate_teacher_forced_sequence

    def record(self, seq):
        seen.append(list(seq))
        return orig(self, seq)
    monkeypatch.setattr(mi.TeacherForceTask, "validate_teacher_forced_sequence", record)
    return seen


@pytest.mark.parametrize("sc", list(GENERATE_SCENARIOS))
def test_generate_sequence_matches_reference(sc, raw_sequences):
    z = np.load(os.path.join(GOLDEN, "generate_greedy.npz"))
    task = _task(z, sc, 1)
    meta = [int(t) for t in z["meta"]]
    with torch.no_grad():
        seq, mems = task.init_seq_and_mems(meta, len(meta))
        verdict = task.generate_sequence(seq, mems)
    assert raw_sequences[-1] == z[sc + "/raw"].tolist()
    assert (verdict is not None) == bool(z[sc + "/valid"][0])


@pytest.mark.parametrize("sc", list(GENERATE_SCENARIOS))
def test_generate_batch_matches_reference(sc, raw_sequences):
    """Three sequences decoded together (one token step + one sampler launch per step for all of them): at
    temperature 0 each must equal the reference's sequence."""
    z = np.load(os.path.join(GOLDEN, "generate_greedy.npz"))
    task = _task(z, sc, 3)
    out = task.generate_batch([int(t) for t in z["meta"]], 3)
    assert len(raw_sequences) == 3
    for s in raw_sequences:
        assert s == z[sc + "/raw"].tolist()
    assert all((o is not None) == bool(z[sc + "/valid"][0]) for o in out)
