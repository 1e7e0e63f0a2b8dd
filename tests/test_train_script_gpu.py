"""End-to-end run of the drop-in train.py on a synthetic corpus (single GPU): train steps, the
same_length evaluation path, checkpoints in the reference format (pickled BaseVocab), resume."""
import glob
import importlib.util
import os
import tempfile

import pytest
import torch

from helpers import ROOT

pytestmark = pytest.mark.gpu

OPTS = ("MODEL.num_layers=2,MODEL.num_heads=2,MODEL.units=128,MODEL.inner_size=256,MODEL.dropout=0.0,"
        "MODEL.attention_dropout=0.0,TRAIN.batch_size=8,TRAIN.batch_chunk=2,TRAIN.tgt_length=64,TRAIN.mem_length=64,"
        "TRAIN.log_interval=2,TRAIN.eval_interval=3,TRAIN.warmup_step=2,EVALUATE.batch_size=4,EVALUATE.tgt_length=32,"
        "EVALUATE.mem_length=64")


def _train_module():
    spec = importlib.util.spec_from_file_location("commu_train_entry", os.path.join(ROOT, "train.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_train_eval_checkpoint_resume():
    from commu.model.dataset import write_synthetic_dataset
    tr = _train_module()
    data = tempfile.mkdtemp()
    work = tempfile.mkdtemp()
    write_synthetic_dataset(data, n_train=40, n_val=6, length=200, ragged=True)
    os.environ.pop("WORLD_SIZE", None)
    tr.main(["--data_dir", data, "--work_dir", work, "--opts", OPTS, "--max_step", "6"])
    ckpts = sorted(glob.glob(os.path.join(work, "*", "checkpoint_*.pt")))
    assert any(c.endswith("checkpoint_last.pt") for c in ckpts) and any(c.endswith("checkpoint_best.pt") for c in ckpts)
    ck = torch.load(ckpts[0], map_location="cpu", weights_only=False)
    assert set(ck) >= {"model", "optimizer", "train_step", "scheduler", "best_val_loss", "vocab", "amp"}
    assert ck["amp"] is None and len(ck["vocab"]) == 729 and "layers.0.dec_attn.qkv_net.weight" in ck["model"]
    assert all(torch.isfinite(v).all() for v in ck["model"].values() if v.is_floating_point())
    # resume and continue two more steps
    last = [c for c in ckpts if c.endswith("checkpoint_last.pt")][0]
    work2 = tempfile.mkdtemp()
    tr.main(["--data_dir", data, "--work_dir", work2, "--opts", OPTS, "--max_step", "8", "--resume", last])
