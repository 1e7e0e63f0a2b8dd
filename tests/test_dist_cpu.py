"""World-size-2 gloo tests (CPU) of the N>1 host logic: the single sum all-reduce of the flat
gradient + 1/world scaling reproduces DDP's averaged gradients (train.py:155, 467-473), per-rank
data streams differ, evaluation shards cover the split exactly once."""
import os
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import orc


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = orc.make_cfg(n_layer=1, n_head=2, d_model=16, d_inner=32, tgt_len=6, mem_len=6, n_token=23)
    P = orc.init_params(cfg, seed=1, std=0.1)
    g = torch.Generator().manual_seed(100)
    data = torch.randint(1, 23, (6, 4), generator=g)
    target = torch.randint(1, 23, (6, 4), generator=g)
    # this rank's shard of the global batch
    cols = slice(rank * 2, rank * 2 + 2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in P.items()}
    nll, _ = orc.forward_loss(cfg, leaves, data[:, cols], target[:, cols], None, None)
    nll.mean().backward()
    names = list(P.keys())
    flat = torch.cat([leaves[k].grad.reshape(-1) for k in names])
    dist.all_reduce(flat)                      # ONE sum all-reduce of the flat arena
    flat /= world                              # folded into the Adam kernel in the product
    if rank == 0:
        full = {k: v.clone().requires_grad_(True) for k, v in P.items()}
        nll_f, _ = orc.forward_loss(cfg, full, data, target, None, None)
        nll_f.mean().backward()                # equal shard sizes: mean over all = mean of shard means
        ref = torch.cat([full[k].grad.reshape(-1) for k in names])
        torch.save({"err": float((flat - ref).abs().max()), "scale": float(ref.abs().max())},
                   os.path.join(tmp, "res.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_allreduce_equals_ddp_average():
    tmp = tempfile.mkdtemp()
    port = 29000 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
    r = torch.load(os.path.join(tmp, "res.pt"))
    assert r["err"] < 1e-6 * max(1.0, r["scale"]), r


def test_rank_streams_and_eval_shards():
    from commu.model.dataset import ComMUDataset, write_synthetic_dataset
    d = tempfile.mkdtemp()
    write_synthetic_dataset(d, 30, 9, 40, ragged=True)
    ds = ComMUDataset(d, None, verbose=False)
    a = next(ds.get_iterator(4, 8, "cpu", "train", True, seed=1111)())[0]
    b = next(ds.get_iterator(4, 8, "cpu", "train", True, seed=2111)())[0]
    assert not torch.equal(a, b)
    tok = [sum(n for *_, n in ds.eval_iterator(2, 8, "cpu", "valid", r, 2)()) for r in range(2)]
    assert sum(tok) == int((ds.valid_seq_length - 1).sum())
