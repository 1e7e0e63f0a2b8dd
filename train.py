"""Training entry point with the reference's command line (train.py of POZAlabs/ComMU-code:
`--data_dir`, `--work_dir`, `--local_rank`), checkpoint format and log lines, driving the native
sm_100a engine:

    python train.py --data_dir D --work_dir W                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train.py --data_dir D --work_dir W

Differences from the reference driver (all additive): `--local-rank` / LOCAL_RANK are accepted too
(torch >= 2.0 launchers), `--opts SECTION.field=value,...` (or COMMU_CFG_OPTS) overrides config
fields, `--max_step N` shortens a run, nothing executes at import time.  Data parallelism does ONE
NCCL gradient all-reduce per optimizer step from the native library instead of DDP's per-micro-batch
bucket all-reduces (same mathematics; see commu-code_b200/commu/engine/trainer.py).
"""
import argparse
import ast
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from commu.engine.trainer import GradComm, Trainer  # noqa: E402
from commu.model.config_helper import get_default_cfg_training  # noqa: E402
from commu.model.dataset import ComMUDataset  # noqa: E402
from commu.model.exp_utils import logging_config  # noqa: E402
from commu.model.model import MemTransformerLM  # noqa: E402
from logger import logger  # noqa: E402


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="ComMU Transformer-XL training (B200-native)")
    ap.add_argument("--data_dir", type=str, required=True, help="location of the data corpus")
    ap.add_argument("--local_rank", "--local-rank", type=int, default=int(os.environ.get("LOCAL_RANK", 0)))
    ap.add_argument("--work_dir", type=str, required=True, help="Base directory to save the trained model.")
    ap.add_argument("--opts", type=str, default="", help="config overrides: MODEL.num_layers=12,TRAIN.lr=0.001")
    ap.add_argument("--max_step", type=int, default=None)
    ap.add_argument("--resume", type=str, default=None,
                    help="checkpoint (.pt written by this script or by the reference) to continue from: model, "
                         "Adam moments and step counter are restored (the reference only ever saves)")
    return ap.parse_args(argv)


def init_model_weights(model, cfg):
    """Reference recipe (train.py:291-342): N(0, base_init) matrices / embedding / r_*_bias, zero
    biases, LayerNorm gain N(1, base_init)."""
    std = cfg.INITIALIZER.base_init
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("layer_norm.weight"):
                p.normal_(1.0, std)
            elif name.endswith(".bias") and p.dim() == 1:
                p.zero_()
            else:
                p.normal_(0.0, std)


def evaluate(model, cfg, eval_iter, pad_id):
    """Reference evaluate() (train.py:74-110): eval lengths, same_length=True, memory carried across
    the segments of a sample group."""
    model.eval()
    model.reset_length(cfg.EVALUATE.tgt_length, cfg.EVALUATE.mem_length)
    model.same_length = True
    tot_tok, tot_nll = 0, 0.0
    with torch.no_grad():
        mems = None
        for data, target, all_reset, n_tok in eval_iter():
            if all_reset:
                mems = None
            loss, mems = model(data, target, None, mems)
            loss = loss[target != pad_id].mean()
            tot_nll += n_tok * loss.float().item()
            tot_tok += n_tok
    model.reset_length(cfg.TRAIN.tgt_length, cfg.TRAIN.mem_length)
    model.same_length = cfg.MODEL.same_length
    model.train()
    return tot_tok, tot_nll


def save_checkpoint(work_dir, rank, distributed, model, trainer, vocab, train_step, best_val, name):
    ckpt = {"model": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()},
            "optimizer": trainer.optimizer_state_dict(), "train_step": train_step,
            "scheduler": {"last_epoch": train_step, "base_lrs": [trainer.base_lr]},
            "best_val_loss": best_val, "vocab": vocab, "amp": None}
    path = os.path.join(work_dir, name)
    logger.info(f"Saving checkpoint to {path}")
    if rank == 0:
        torch.save(ckpt, path)
    if distributed:
        dist.barrier()


def resume_from(path, model, trainer, device):
    """Continues from a checkpoint written by this script OR by the reference's save_checkpoint (train.py:29-54 of the
    reference: {"model", "optimizer", "train_step", "scheduler", "best_val_loss", "vocab", "amp"}): weights with
    strict=False like the reference's loader (model_initializer.py:46-47), torch.optim.Adam's exp_avg / exp_avg_sq /
    step per parameter index, the step counter that drives the LambdaLR schedule.  Returns (train_step, best_val)."""
    ckpt = torch.load(path, map_location=device, weights_only=False)      # pickled BaseVocab inside
    model.load_state_dict(ckpt["model"], strict=False)
    if ckpt.get("optimizer"):
        trainer.load_optimizer_state_dict(ckpt["optimizer"])
    train_step = int(ckpt.get("train_step", 0))
    trainer.step = train_step
    trainer.engine.refresh_shadow()
    return train_step, ckpt.get("best_val_loss")


def reduce_scalars(vals, device, distributed):
    t = torch.tensor(vals, dtype=torch.float64, device=device)
    if distributed:
        dist.all_reduce(t)
    return t.tolist()


def main(argv=None):
    args = parse_args(argv)
    overrides = {}
    for item in filter(None, (s.strip() for s in args.opts.split(","))):
        k, v = item.split("=", 1)
        try:
            overrides[k] = ast.literal_eval(v)
        except (ValueError, SyntaxError):
            overrides[k] = v
    if args.max_step:
        overrides["TRAIN.max_step"] = args.max_step
    cfg = get_default_cfg_training(overrides)
    torch.cuda.set_device(args.local_rank)
    device = torch.device("cuda", args.local_rank)
    distributed = int(os.environ.get("WORLD_SIZE", "1")) > 1
    if distributed:
        dist.init_process_group(backend="nccl", init_method="env://", device_id=device)
    world = dist.get_world_size() if distributed else 1
    rank = dist.get_rank() if distributed else 0

    stamp = torch.tensor(time.time(), dtype=torch.float64, device=device)
    if distributed:
        dist.broadcast(stamp, 0)
    work_dir = os.path.join(args.work_dir, time.strftime("%Y%m%d-%H%M%S", time.localtime(float(stamp))))
    os.makedirs(work_dir, exist_ok=True)
    if rank == 0:
        with open(os.path.join(work_dir, "config.yml"), "w") as f:
            f.write(str(cfg))
    logging_config(work_dir, "train_rank{}".format(rank), console=(rank == 0))

    seed = cfg.TRAIN.seed
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)

    logger.info("Loading data")
    dataset = ComMUDataset(args.data_dir, cfg, verbose=(rank == 0))
    vocab = dataset.vocab
    assert cfg.TRAIN.batch_size % world == 0
    batch_size = cfg.TRAIN.batch_size // world
    assert batch_size % cfg.TRAIN.batch_chunk == 0
    train_iter = dataset.get_iterator(batch_size, cfg.TRAIN.tgt_length, device, "train", True,
                                      seed=seed + rank * 1000)
    val_iter = dataset.eval_iterator(cfg.EVALUATE.batch_size, cfg.EVALUATE.tgt_length, device, "valid",
                                     local_rank=rank, world_size=world)
    test_iter = dataset.eval_iterator(cfg.EVALUATE.batch_size, cfg.EVALUATE.tgt_length, device, "test",
                                      local_rank=rank, world_size=world)

    logger.info("Build the model")
    assert cfg.MODEL.units % cfg.MODEL.num_heads == 0
    model = MemTransformerLM(cfg, vocab)
    init_model_weights(model, cfg)
    n_all = sum(p.nelement() for p in model.parameters())
    n_nonemb = sum(p.nelement() for p in model.layers.parameters())
    model = model.to(device)
    model.train()
    if distributed:                       # same initial weights on every rank (DDP did this broadcast)
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    comm = GradComm(rank, world, device) if distributed else None
    trainer = Trainer(model, lr=cfg.TRAIN.lr / world, warmup_step=cfg.TRAIN.warmup_step, lr_min=cfg.TRAIN.lr_min,
                      clip=cfg.TRAIN.clip, batch_chunk=cfg.TRAIN.batch_chunk, pad_id=vocab.pad_id, world=world,
                      comm=comm, global_lr=cfg.TRAIN.lr, weight_decay=cfg.TRAIN.weight_decay)
    logger.info("=" * 100)
    logger.info(args)
    logger.info("=" * 100)
    logger.info("#total params = {}".format(n_all))
    logger.info("#non emb params in generator = {}".format(n_nonemb))
    logger.info("Start training")

    train_step, best_val = 0, np.inf
    if args.resume:
        train_step, bv = resume_from(args.resume, model, trainer, device)
        if bv is not None:
            best_val = bv
        logger.info("Resumed from {} at step {}".format(args.resume, train_step))
    log_loss = torch.zeros((), device=device, dtype=torch.float64)
    log_gnorm = torch.zeros((), device=device, dtype=torch.float64)
    log_tok = 0
    t_log = time.time()
    for data, target, reset, n_tok in train_iter():
        loss, gnorm = trainer.train_step(data, target, reset)
        train_step += 1
        # the reference logs sum_i loss_i * count_i * chunks; with equal chunk sizes this equals loss * tokens
        log_loss += loss.double() * n_tok
        log_gnorm += gnorm.double()
        log_tok += int(n_tok)

        if train_step % cfg.TRAIN.log_interval == 0:
            tot_loss, tot_gn, tot_tok = reduce_scalars([float(log_loss), float(log_gnorm), log_tok], device, distributed)
            nll = tot_loss / tot_tok
            if rank == 0:
                logger.info("Train Step {}/{}, lr={:f}, tokens/s={:.1f}, nll={:.4f}, ppl={:.2f}, grad norm={}, ".format(
                    train_step, cfg.TRAIN.max_step, trainer.current_lr(), tot_tok / (time.time() - t_log), nll,
                    math.exp(nll), tot_gn / (cfg.TRAIN.log_interval * world)))
            log_loss.zero_()
            log_gnorm.zero_()
            log_tok = 0
            t_log = time.time()

        if train_step % cfg.TRAIN.eval_interval == 0:
            t0 = time.time()
            v_tok, v_nll = evaluate(model, cfg, val_iter, vocab.pad_id)
            v_tok, v_nll = reduce_scalars([v_tok, v_nll], device, distributed)
            val_nll = v_nll / v_tok
            if rank == 0:
                logger.info("Eval step {}, time={}s, val nll={}, val ppl={},".format(
                    train_step, time.time() - t0, val_nll, math.exp(val_nll)))
            save_checkpoint(work_dir, rank, distributed, model, trainer, vocab, train_step, val_nll, "checkpoint_last.pt")
            if not best_val or val_nll < best_val:
                best_val = val_nll
                save_checkpoint(work_dir, rank, distributed, model, trainer, vocab, train_step, best_val,
                                "checkpoint_best.pt")
                t0 = time.time()
                t_tok, t_nll = evaluate(model, cfg, test_iter, vocab.pad_id)
                t_tok, t_nll = reduce_scalars([t_tok, t_nll], device, distributed)
                if rank == 0:
                    logger.info("Test step {}, time={}s, test nll={}, test ppl={}, #evaluated tokens={}".format(
                        train_step, time.time() - t0, t_nll / t_tok, math.exp(t_nll / t_tok), t_tok))

        if train_step == cfg.TRAIN.max_step:
            logger.info("-" * 100)
            logger.info("End of training")
            break

    best_path = os.path.join(work_dir, "checkpoint_best.pt")
    if os.path.exists(best_path):
        ckpt = torch.load(best_path, map_location=device, weights_only=False)
        model.load_state_dict(ckpt["model"])
        model._engine().refresh_shadow()
        t_tok, t_nll = evaluate(model, cfg, test_iter, vocab.pad_id)
        t_tok, t_nll = reduce_scalars([t_tok, t_nll], device, distributed)
        logger.info("=" * 100)
        logger.info("| End of training | test nll {:5.2f} | test ppl {:9.3f}".format(t_nll / t_tok, math.exp(t_nll / t_tok)))
        logger.info("=" * 100)
    if comm is not None:
        comm.close()
    if distributed:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
