"""Shared helpers for tests: golden loading and the oracle import (test-only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import transfoxl_oracle as orc  # noqa: E402


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    c = z["cfg"]
    cfg = orc.make_cfg(n_layer=int(c[0]), n_head=int(c[1]), d_model=int(c[2]), d_inner=int(c[3]),
                       tgt_len=int(c[4]), mem_len=int(c[5]), same_length=bool(c[6]),
                       clamp_len=int(c[7]), n_token=int(c[8]))
    params = {k[len("param/"):]: torch.from_numpy(z[k]) for k in z.files if k.startswith("param/")}
    params.pop("pos_emb.inv_freq", None)
    return z, cfg, params


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
