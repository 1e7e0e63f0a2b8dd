"""Reference-saved checkpoint round trip at the reference's code-default model shape (6 layers, 10 heads of 50,
d_model 500, d_inner 1000: head dim and widths that are NOT multiples of 64, so every padded layout is exercised).

tests/golden/make_golden.py::gen_checkpoint_case had the UNMODIFIED reference write such a checkpoint with torch.save
(model.state_dict() incl. the tied crit weight and the inv_freq buffer, torch.optim.Adam.state_dict(),
LambdaLR.state_dict(), train_step, pickled BaseVocab), re-read it and continue training for three optimizer steps; the
159 MB file cannot be a fixture, so its content is a pure function of the key names and the fixture keeps its structure
plus the reference's continued losses / learning rates / gradient norms.  This test rebuilds the identical file, resumes
from it through train.py's resume_from and must reproduce the reference's continuation (loss within 1e-3 relative)."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT

pytestmark = pytest.mark.gpu

sys.path.insert(0, GOLDEN)
from make_golden import CKPT_CFG, CKPT_HYPER, ckpt_batches, keyed_tensor  # noqa: E402


def _rebuild_reference_checkpoint(path, st):
    from commu.model.dataset import BaseVocab
    model = {}
    for k, shape, dtype in st["model_keys"]:
        if k == "crit.out_layers.0.weight":
            continue
        if k == "pos_emb.inv_freq":
            model[k] = 1.0 / (10000 ** (torch.arange(0.0, CKPT_CFG["d_model"], 2.0) / CKPT_CFG["d_model"]))
            continue
        kind = "ln_weight" if k.endswith("layer_norm.weight") else "w"
        model[k] = keyed_tensor("model/" + k, tuple(shape), kind)
    model["crit.out_layers.0.weight"] = model["word_emb.emb_layers.0.weight"]      # tied storage, like the reference
    model = {k: model[k] for k, _, _ in st["model_keys"]}                           # the reference's key order
    shapes = {k: tuple(s) for k, s, _ in st["model_keys"]}
    state = {i: {"step": torch.tensor(float(CKPT_HYPER["train_step"])),
                 "exp_avg": keyed_tensor("opt/exp_avg/" + k, shapes[k], "exp_avg"),
                 "exp_avg_sq": keyed_tensor("opt/exp_avg_sq/" + k, shapes[k], "exp_avg_sq")}
             for i, k in enumerate(st["param_order"])}
    ckpt = {"model": model, "optimizer": {"state": state, "param_groups": st["optimizer_param_groups"]},
            "train_step": CKPT_HYPER["train_step"],
            "scheduler": {"last_epoch": CKPT_HYPER["train_step"], "base_lrs": [CKPT_HYPER["lr"]]},
            "best_val_loss": 3.21, "vocab": BaseVocab(), "amp": None}
    assert list(ckpt.keys()) == st["top_level_keys"]
    assert st["vocab_class"] == "commu.model.dataset.BaseVocab"
    torch.save(ckpt, path)


def test_resume_reference_checkpoint_default_shape(tmp_path):
    from types import SimpleNamespace as NS
    from commu.engine.trainer import Trainer
    from commu.model.dataset import BaseVocab
    from commu.model.model import MemTransformerLM
    sys.path.insert(0, ROOT)
    import train as train_script
    z = np.load(os.path.join(GOLDEN, "checkpoint_default_shape.npz"))
    st = json.loads(str(z["structure"]))
    h, c = CKPT_HYPER, CKPT_CFG
    path = str(tmp_path / "checkpoint_last.pt")
    _rebuild_reference_checkpoint(path, st)
    cfg = NS(MODEL=NS(num_layers=c["n_layer"], num_heads=c["n_head"], units=c["d_model"], inner_size=c["d_inner"],
                      dropout=0.0, attention_dropout=0.0, same_length=False, clamp_len=-1),
             TRAIN=NS(tgt_length=c["tgt_len"], mem_length=c["mem_len"]))
    torch.manual_seed(5)
    model = MemTransformerLM(cfg, BaseVocab()).cuda()
    # every reference key is known to the drop-in module (strict load would pass as well)
    missing = set(k for k, _, _ in st["model_keys"]) ^ set(model.state_dict().keys())
    assert not missing, missing
    model.train()
    tr = Trainer(model, lr=h["lr"], warmup_step=h["warmup"], lr_min=h["lr_min"], clip=1.0, batch_chunk=h["chunks"])
    step, best = train_script.resume_from(path, model, tr, torch.device("cuda"))
    assert step == h["train_step"] and abs(best - 3.21) < 1e-12
    sd = tr.optimizer_state_dict()
    assert sorted(sd["state"][0].keys()) == st["optimizer_state_keys"] and len(sd["state"]) == st["optimizer_n_state"]
    for s, (data, target, reset) in enumerate(ckpt_batches(h["n_steps"], c["tgt_len"], h["B"], h["data_seed"])):
        assert abs(tr.current_lr() - z["lrs"][s]) < 1e-12, s
        loss, gn = tr.train_step(data.cuda(), target.cuda(), reset.cuda())
        assert abs(float(loss) - z["losses"][s]) / z["losses"][s] < 1e-3, (s, float(loss), z["losses"][s])
        assert abs(float(gn) - z["gnorms"][s]) / z["gnorms"][s] < 0.03, (s, float(gn), z["gnorms"][s])
    # and the file this repo writes is read back by the same path (save -> resume -> identical next loss)
    train_script.save_checkpoint(str(tmp_path), 0, False, model, tr, BaseVocab(), step + h["n_steps"], 3.0, "again.pt")
    data, target, reset = ckpt_batches(h["n_steps"] + 1, c["tgt_len"], h["B"], h["data_seed"])[-1]
    mems_before = list(tr.mems)
    l_a, _ = tr.train_step(data.cuda(), target.cuda(), reset.cuda())
    model2 = MemTransformerLM(cfg, BaseVocab()).cuda()
    model2.train()
    tr2 = Trainer(model2, lr=h["lr"], warmup_step=h["warmup"], lr_min=h["lr_min"], clip=1.0, batch_chunk=h["chunks"])
    step2, _ = train_script.resume_from(str(tmp_path / "again.pt"), model2, tr2, torch.device("cuda"))
    assert step2 == step + h["n_steps"]
    tr2.mems = mems_before                              # the recurrent memory is run state, not checkpoint state
    l_b, _ = tr2.train_step(data.cuda(), target.cuda(), reset.cuda())
    assert abs(float(l_a) - float(l_b)) < 1e-6 * abs(float(l_a)) + 1e-6, (float(l_a), float(l_b))
