"""CPU tests of the host-side logic: dataset batching vs batches produced by the reference class
(golden), config shim, LR schedule, C-ABI library loading / exported symbols, loud failure without
CUDA."""
import ctypes
import os
import re
import tempfile

import numpy as np
import pytest
import torch

from helpers import GOLDEN, ROOT


def _write_corpus(z, d):
    for tag in ("train", "val"):
        idx = sorted({int(k.split("/")[2]) for k in z.files if k.startswith("corpus/%s/" % tag)})
        inp = np.empty(len(idx), dtype=object)
        tgt = np.empty(len(idx), dtype=object)
        for i in idx:
            inp[i] = z["corpus/%s/%d/input" % (tag, i)]
            tgt[i] = z["corpus/%s/%d/target" % (tag, i)]
        np.save(os.path.join(d, "input_%s.npy" % tag), inp, allow_pickle=True)
        np.save(os.path.join(d, "target_%s.npy" % tag), tgt, allow_pickle=True)


def test_dataset_batches_match_reference():
    from commu.model.dataset import ComMUDataset
    z = np.load(os.path.join(GOLDEN, "dataset_batches.npz"), allow_pickle=True)
    d = tempfile.mkdtemp()
    _write_corpus(z, d)
    ds = ComMUDataset(d, None, verbose=False)
    it = ds.get_iterator(5, 16, "cpu", "train", True, seed=1111)()
    for b in range(40):
        data, target, reset, ntok = next(it)
        assert np.array_equal(data.numpy(), z["train/%d/data" % b]), b
        assert np.array_equal(target.numpy(), z["train/%d/target" % b]), b
        assert np.array_equal(reset.numpy(), z["train/%d/reset" % b]), b
        assert ntok == int(z["train/%d/ntok" % b])
    for rank in range(2):
        batches = list(ds.eval_iterator(3, 16, "cpu", "valid", rank, 2)())
        n_ref = len([k for k in z.files if k.startswith("eval%d/" % rank) and k.endswith("/data")])
        assert len(batches) == n_ref
        for b, (data, target, first, ntok) in enumerate(batches):
            assert np.array_equal(data.numpy(), z["eval%d/%d/data" % (rank, b)])
            assert np.array_equal(target.numpy(), z["eval%d/%d/target" % (rank, b)])
            assert bool(first) == bool(z["eval%d/%d/first" % (rank, b)])
            assert ntok == int(z["eval%d/%d/ntok" % (rank, b)])


def test_config_defaults_and_overrides():
    from commu.model.config_helper import get_default_cfg_training, get_default_cfg_inference
    c = get_default_cfg_training()
    assert (c.MODEL.num_layers, c.MODEL.num_heads, c.MODEL.units, c.MODEL.inner_size) == (6, 10, 500, 1000)
    assert (c.TRAIN.tgt_length, c.TRAIN.mem_length, c.TRAIN.batch_chunk, c.TRAIN.lr) == (128, 1024, 4, 0.004)
    assert c.EVALUATE.mem_length == 2048 and c.INITIALIZER.base_init == 0.01
    with pytest.raises(AttributeError):
        c.MODEL.units = 1
    c.defrost(); c.MODEL.same_length = True; c.freeze()
    assert "same_length: True" in str(c)
    c2 = get_default_cfg_training({"MODEL.num_layers": 12, "TRAIN.tgt_length": 2048})
    assert c2.MODEL.num_layers == 12 and c2.TRAIN.tgt_length == 2048
    assert get_default_cfg_inference().MODEL.memory_length == 4146


def test_lr_schedule():
    from commu.engine.trainer import lr_multiplier
    assert lr_multiplier(0, 100, 0.004, 1e-4) == 0.0
    assert lr_multiplier(50, 100, 0.004, 1e-4) == 0.5
    assert abs(lr_multiplier(400, 100, 0.004, 1e-4) - 0.5) < 1e-12
    assert lr_multiplier(10 ** 9, 100, 0.004, 1e-4) == 1e-4 / 0.004


def test_library_exports_every_declared_symbol():
    from commu import _native as nv
    hdr = open(os.path.join(ROOT, "include", "commu_b200.h")).read()
    declared = set(re.findall(r"\b(commu_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(nv.LIB_PATH)
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    # and every bound signature refers to a declared symbol
    for sym in nv.SIGNATURES:
        assert sym in declared, sym
    assert lib.commu_abi_version() == 3


def test_no_cpu_fallback():
    from types import SimpleNamespace as NS
    from commu.model.model import MemTransformerLM

    class V:
        def __len__(self):
            return 50
    cfg = NS(MODEL=NS(num_layers=1, num_heads=2, units=32, inner_size=64, dropout=0.0, attention_dropout=0.0,
                      same_length=False, clamp_len=-1), TRAIN=NS(tgt_length=4, mem_length=4))
    m = MemTransformerLM(cfg, V())
    x = torch.zeros(4, 1, dtype=torch.long)
    with pytest.raises(RuntimeError):
        m(x, x, None, None)
    with pytest.raises(RuntimeError):
        m.forward_generate(x, None)


def test_c_structs_match_ctypes_layout(tmp_path):
    """The argument structs of the C-ABI (include/commu_b200.h) and their ctypes mirrors (commu/_native.py) must agree
    field by field: a C program built from the header prints sizeof / offsetof, compared with ctypes."""
    import shutil
    import subprocess
    from commu import _native as nv
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    pairs = {"commu_gemm_args": nv.GemmArgs, "CommuDecLinear": nv.DecLinearArgs}
    lines = []
    for cname, cls in pairs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "commu_b200.h"\nint main(void) {\n%s\nreturn 0; }\n'
                   % "\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for ln in out:
        if not ln:
            continue
        cname, fname, val = ln.split()
        cls = pairs[cname]
        if fname == "sizeof":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, fname).offset == int(val), (cname, fname, getattr(cls, fname).offset, val)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in pairs.values())
