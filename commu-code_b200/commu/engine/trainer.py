"""Native train step: micro-batch forward/backward into a flat fp32 gradient arena, ONE NCCL sum
all-reduce per optimizer step, fused grad-norm -> clip -> Adam, bf16 shadow refresh.

Restates the step tail of the reference (train.py:133-169: batch_chunk loop, pad-masked mean loss,
clip_grad_norm_(1.0), Adam with lr/num_gpus, LambdaLR warm-up / inverse-sqrt schedule :448-461)
with all arithmetic in libcommu_b200.so.  Python keeps only the schedule and the call order.
"""
import ctypes
import glob
import os

import torch

from commu import _native as nv


def lr_multiplier(step, warmup_step, lr, lr_min):
    """train.py:448-461"""
    if step == 0 and warmup_step == 0:
        return 1.0
    if step > warmup_step:
        return max((warmup_step ** 0.5) / (step ** 0.5), lr_min / lr)
    return step / warmup_step


def _find_nccl():
    """The libnccl torch has already mapped into this process (so that both talk to ONE NCCL), else the bundled /
    system copy."""
    try:
        with open("/proc/self/maps") as f:
            for line in f:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    for base in (os.path.dirname(torch.__file__) + "/../nvidia/nccl/lib", "/usr/lib/x86_64-linux-gnu"):
        hits = sorted(glob.glob(os.path.join(base, "libnccl.so*")))
        if hits:
            return os.path.abspath(hits[0])
    return ""


class GradComm:
    """Direct-NCCL communicator of the native library, bootstrapped through torch.distributed
    (the unique id is broadcast with whatever backend the process group already has)."""

    def __init__(self, rank, world, device):
        import torch.distributed as dist
        self.rank, self.world = rank, world
        L = nv.lib()
        path = _find_nccl().encode()
        idbuf = (ctypes.c_char * 128)()
        if rank == 0:
            nv.check(L.commu_comm_unique_id(path, idbuf))
        t = torch.frombuffer(bytearray(bytes(idbuf)), dtype=torch.uint8).clone()
        if dist.get_backend() == "nccl":
            t = t.to(device)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().numpy().tobytes())
        nv.check(L.commu_comm_init(path, ctypes.c_char_p(raw), rank, world))

    def allreduce_(self, flat):
        L = nv.lib()
        L.commu_allreduce_sum_f32.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        nv.check(L.commu_allreduce_sum_f32(ctypes.c_void_p(flat.data_ptr()), flat.numel(), nv.stream_ptr()))

    def close(self):
        nv.check(nv.lib().commu_comm_destroy())


class Trainer:
    def __init__(self, model, lr, warmup_step=0, lr_min=0.0, clip=1.0, betas=(0.9, 0.999), eps=1e-8,
                 batch_chunk=1, pad_id=0, world=1, comm=None, global_lr=None, weight_decay=0.0):
        """lr is the per-rank learning rate (reference: cfg.TRAIN.lr / num_gpus, train.py:441); the floor of the
        schedule is the ratio lr_min / global_lr with the UNDIVIDED cfg.TRAIN.lr (train.py:456), so callers that divide
        lr by the world size pass global_lr = cfg.TRAIN.lr (default: lr, i.e. one rank).  weight_decay is Adam's L2
        term (g += wd * p before the moment updates, torch.optim.Adam; train.py:442-443)."""
        self.model = model
        self.base_lr, self.warmup_step, self.lr_min = lr, warmup_step, lr_min
        self.global_lr = lr if global_lr is None else global_lr
        self.weight_decay = float(weight_decay)
        self.clip, self.betas, self.eps = clip, betas, eps
        self.batch_chunk, self.pad_id = batch_chunk, pad_id
        self.world, self.comm = world, comm
        self.step = 0
        self.mems = [None] * batch_chunk
        self._flatten()

    # fp32 master weights, gradients and Adam moments live in flat arenas; every nn.Parameter (and
    # its .grad) is a view, so state_dict() / checkpoints keep the reference names and shapes.
    @torch.no_grad()
    def _flatten(self):
        plist = self.model._params()
        names = self.model._param_names
        dev = plist[0].device
        offs, total = [], 0
        for p in plist:
            offs.append(total)
            total += (p.numel() + 63) // 64 * 64
        self.flat_p = torch.zeros(total, device=dev)
        self.flat_g = torch.zeros(total, device=dev)
        self.flat_m = torch.zeros(total, device=dev)
        self.flat_v = torch.zeros(total, device=dev)
        self.grads = {}
        for n, p, o in zip(names, plist, offs):
            view = self.flat_p[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            g = self.flat_g[o:o + p.numel()].view_as(p)
            p.grad = g
            self.grads[n] = g
        self.offsets = dict(zip(names, offs))
        # [lo, hi) of every decoder layer inside the flat arenas (a layer's parameters are consecutive), and the rest
        self.layer_spans = []
        for l in range(self.model.n_layer):
            idx = [i for i, n in enumerate(names) if n.startswith("layers.%d." % l)]
            lo, last = offs[idx[0]], idx[-1]
            hi = offs[last + 1] if last + 1 < len(offs) else total
            assert idx == list(range(idx[0], last + 1))
            self.layer_spans.append((lo, hi))
        covered = sorted(self.layer_spans)
        self.other_spans, pos = [], 0
        for lo, hi in covered + [(total, total)]:
            if lo > pos:
                self.other_spans.append((pos, lo))
            pos = max(pos, hi)
        self.side = torch.cuda.Stream(dev) if dev.type == "cuda" else None
        self.gnorm_sq = torch.zeros(1, device=dev)
        self.gnorm = torch.zeros(1, device=dev)
        self.engine = self.model._engine()
        # aligned shapes: the bf16 GEMM operands are views of ONE flat bf16 arena with the layout of flat_p, written by
        # the Adam kernel itself (no cast pass after the optimizer step)
        self.flat_pb = None
        if getattr(self.engine, "aligned", False) and dev.type == "cuda":
            self.flat_pb = self.flat_p.to(torch.bfloat16)
            self.engine.adopt_shadow(self.flat_pb, self.flat_p, self.offsets)
        self.engine.refresh_shadow()

    def current_lr(self):
        return self.base_lr * lr_multiplier(self.step, self.warmup_step, self.global_lr, self.lr_min)

    @torch.no_grad()
    def train_step(self, data, target, reset):
        """data/target: int64 [T, B] device tensors, reset: bool [B] (or None).
        Returns (loss, grad_norm) as 0-d device tensors: loss = sum over chunks of the reference's
        `loss[target != pad].mean() / batch_chunk` (train.py:148-149)."""
        total = self.accumulate_gradients(data, target, reset)
        return total, self.apply_update()

    @torch.no_grad()
    def accumulate_gradients(self, data, target, reset, exchange=True):
        """Forward / backward of every micro-batch into the flat gradient arena; with `exchange` (and a communicator)
        the arena then holds the SUM over ranks (the 1 / world factor is folded into apply_update)."""
        m = self.model
        eng = self.engine
        C = self.batch_chunk
        self.flat_g.zero_()
        dcs = torch.chunk(data, C, 1)
        tcs = torch.chunk(target, C, 1)
        rcs = torch.chunk(reset, C, 0) if reset is not None else [None] * C
        total = torch.zeros((), device=data.device)
        dist_on = exchange and self.comm is not None and self.world > 1
        for i in range(C):
            d_i, t_i = dcs[i].contiguous(), tcs[i].contiguous()
            r_i = rcs[i].contiguous() if rcs[i] is not None else None
            nll, self.mems[i] = eng.forward_loss(d_i, t_i, r_i, self.mems[i], m.mem_len, m.same_length,
                                                 m.clamp_len, save=True, dropout=m._dropout_arg())
            mask = (t_i != self.pad_id).float()
            w = mask / (mask.sum() * C)
            total += (nll * w).sum()
            overlap = dist_on and i == C - 1 and self.overlap
            eng.backward(w, self.grads, layer_done=self._exchange_layer if overlap else None)
        if dist_on:
            if self.overlap:
                # the layers' shares went out on the side stream while the backward was still running (layer l's
                # share as soon as its last gradient was issued); what is left is the embedding / bias remainder
                for lo, hi in self.other_spans:
                    self._exchange(lo, hi)
                torch.cuda.current_stream().wait_stream(self.side)
            else:
                self.comm.allreduce_(self.flat_g)
        return total

    @torch.no_grad()
    def apply_update(self):
        """clip_grad_norm_(clip) + Adam on the flat arenas, bf16 shadow refresh; returns the gradient norm."""
        lr = self.current_lr()
        self.step += 1
        self.gnorm_sq.zero_()
        nv.call("commu_sumsq", self.flat_g, self.flat_g.numel(), self.gnorm_sq)
        nv.call("commu_clip_adam", self.flat_p, self.flat_g, self.flat_m, self.flat_v, self.flat_p.numel(),
                lr, self.betas[0], self.betas[1], self.eps, self.step, self.gnorm_sq, self.clip,
                1.0 / self.world, self.weight_decay, self.gnorm, self.flat_pb)
        self.engine.refresh_shadow(after_update=self.flat_pb is not None)
        return self.gnorm[0].clone()

    # ---- gradient exchange overlapped with the backward (one NCCL sum all-reduce per layer span, side stream) ----
    overlap = True

    def _exchange(self, lo, hi):
        ev = torch.cuda.Event()
        ev.record()                                  # everything that wrote flat_g[lo:hi] is in the stream before this
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            self.comm.allreduce_(self.flat_g[lo:hi])

    def _exchange_layer(self, l):
        self._exchange(*self.layer_spans[l])

    def optimizer_state_dict(self):
        """Same structure as torch.optim.Adam.state_dict() so reference-style checkpoints load."""
        state, idx = {}, 0
        for n, p in zip(self.model._param_names, self.model._param_list):
            o = self.offsets[n]
            state[idx] = {"step": torch.tensor(float(self.step)),
                          "exp_avg": self.flat_m[o:o + p.numel()].view_as(p).clone(),
                          "exp_avg_sq": self.flat_v[o:o + p.numel()].view_as(p).clone()}
            idx += 1
        group = {"lr": self.current_lr(), "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "params": list(range(idx))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        for idx, (n, p) in enumerate(zip(self.model._param_names, self.model._param_list)):
            st = sd["state"].get(idx)
            if st is None:
                continue
            o = self.offsets[n]
            self.flat_m[o:o + p.numel()].view_as(p).copy_(st["exp_avg"])
            self.flat_v[o:o + p.numel()].view_as(p).copy_(st["exp_avg_sq"])
            self.step = int(float(st["step"]))
