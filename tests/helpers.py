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


# ---------------------------------------------------------------------------------------------------
# numpy restatement of the kernels' counter-based dropout masks (commu-code_b200/csrc/dropout.cuh):
# the parity tests apply the identical mask to the torch reference.
# ---------------------------------------------------------------------------------------------------
_U32 = np.uint64(0xFFFFFFFF)


def _rand64(ctr, ka, kb):
    """ctr, ka, kb: uint64 arrays holding 32-bit values -> (x_word, y_word) like drop::rand64."""
    c32 = np.uint64(32)
    x = ((ctr ^ ka) & _U32) * np.uint64(0xD2511F53)
    y = ((x >> c32) ^ (x & _U32) ^ kb) & _U32
    z = y * np.uint64(0xCD9E8D57)
    w = ((z >> c32) ^ (z & _U32) ^ ka) & _U32
    r = w * np.uint64(0x9E3779B1)
    return ((r >> c32) ^ y) & _U32, ((r & _U32) ^ (z >> c32)) & _U32


def drop_thr15(p):
    return min(32767, max(0, int(p * 32768.0 + 0.5)))


def drop_keep_prob(p):
    return 1.0 - drop_thr15(p) / 32768.0


def drop_keep_mask(seed, rows, cols, p):
    """bool [rows, cols]: keep decision of element (row, col) under `seed` (cols is rounded up to 4 inside)."""
    ka0, kb0 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    r = np.arange(rows, dtype=np.uint64)[:, None]
    ka = (ka0 ^ ((r * np.uint64(0x9E3779B1)) & _U32)) & _U32
    kb = (kb0 + r * np.uint64(0x85EBCA77)) & _U32
    c4 = np.arange((cols + 3) // 4, dtype=np.uint64)[None, :]
    xw, yw = _rand64(c4, ka, kb)
    m15, c16 = np.uint64(0x7FFF), np.uint64(16)
    f = np.stack([xw & m15, (xw >> c16) & m15, yw & m15, (yw >> c16) & m15], -1).reshape(rows, -1)
    return torch.from_numpy(f[:, :cols] >= np.uint64(drop_thr15(p)))


def attn_keep_mask(seed, B, H, T, K, p):
    """bool [B, H, T, K]: attention-probability keep mask (row id = (b*H + h)*T + i, column = key j)."""
    return drop_keep_mask(seed, B * H * T, K, p).view(B, H, T, K)
