"""Benchmark of the ComMU Transformer-XL training hot path (BASELINE.json configs[1]:
12-layer d512 n_head=8 d_inner=2048, seq_len = mem_len = 2048, vocab 729, synthetic tokens).

    python bench.py --gpus N --steps K --warmup W [--impl native|reference]

One JSON line on stdout (rank 0).  A "step" is one optimizer step (micro-batch forward, backward,
gradient all-reduce for N>1, clip, Adam) over B_PER_GPU synthetic sequences of 2048 tokens per GPU
with the recurrent memory full (2048).  `value` = trained tokens/s of the whole job with inputs
resident in HBM; `e2e` = the same through Trainer.train_step with host (pinned) inputs copied in and
the loss read back every step.  `--impl reference` times the CPU port of the reference algorithm
(oracle/, "kind": "port") on the host cores.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))

import torch  # noqa: E402

CONFIGS = {
    # BASELINE.json configs[1] / [2]: the headline
    "c2": dict(cfg=dict(n_layer=12, n_head=8, d_model=512, d_inner=2048, tgt_len=2048, mem_len=2048, n_token=729),
               batch=16, metric="train tokens/sec @12L d512 seq2048 mem2048",
               workload="BASELINE configs[1]: 12L d512 H8 Di2048 T=2048 M=2048 V=729"),
    # BASELINE.json configs[4]: the scaled stress run (SURVEY.md 8d "C5": B = 2 per GPU)
    "c5": dict(cfg=dict(n_layer=24, n_head=16, d_model=1024, d_inner=4096, tgt_len=4096, mem_len=4096, n_token=729),
               batch=2, metric="train tokens/sec @24L d1024 seq4096 mem4096",
               workload="BASELINE configs[4]: 24L d1024 H16 Di4096 T=4096 M=4096 V=729"),
}
CFG = dict(CONFIGS["c2"]["cfg"])
B_PER_GPU = int(os.environ.get("COMMU_BENCH_BATCH", "16"))
METRIC = CONFIGS["c2"]["metric"]
WORKLOAD = CONFIGS["c2"]["workload"]


def select_config(name):
    global CFG, B_PER_GPU, METRIC, WORKLOAD
    c = CONFIGS[name]
    CFG = dict(c["cfg"])
    B_PER_GPU = int(os.environ.get("COMMU_BENCH_BATCH", str(c["batch"])))
    METRIC, WORKLOAD = c["metric"], c["workload"]


def algorithmic_flops_per_token(L, d, Di, T, M, B, V):
    """minimal-algorithm FLOPs per trained token (causal-visible keys)."""
    K = T + M
    macs_fwd_layer = d * d + 2 * d * d * K / T + d * d * K / (T * B) + 3 * d * (M + (T + 1) / 2) + d * d + 2 * d * Di
    f_fwd = 2 * (L * macs_fwd_layer + d * V)
    f_train = 3 * f_fwd - 2 * L * 2 * d * d * M / T
    attn_fwd_flops_per_bh = 2 * 3 * 64 * (T * M + T * (T + 1) / 2)   # per (b, h), head dim 64: AC + BD + AV
    return f_fwd, f_train, attn_fwd_flops_per_bh


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return dict(bf16=j.get("bf16_tflops_sustained", j.get("bf16_tflops")), bf16_burst=j.get("bf16_tflops"),
                    hbm=j.get("hbm_gbs"), source="measured (MEASURED_PEAKS.json, sustained)")
    return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel class, from the committed `ncu --set full`
    capture (profiles/ncu_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); None when not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p)).get(kernel, {}).get("bytes_per_launch")
    except (OSError, ValueError):
        return None


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        sm, mx, reasons = [], 0.0, set()
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=(sm[len(sm) // 2] if sm else None), sm_max_mhz=mx or None, samples=len(sm),
                    reasons=sorted(reasons))

    def stop(self):
        if self.proc:
            self.proc.kill()


def synthetic_tokens(T, B, n, gen):
    """Full columns: event tokens uniform in [2,559]; no pad, reset rarely true."""
    data = torch.randint(2, 560, (n, T + 1, B), generator=gen, dtype=torch.int64)
    return data[:, :-1].contiguous(), data[:, 1:].contiguous()


def run_native(args):
    import torch.distributed as dist
    from types import SimpleNamespace as NS
    from commu import _native as nv
    from commu.engine.trainer import GradComm, Trainer
    from commu.model.model import MemTransformerLM

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        comm = GradComm(rank, world, dev)

    class Vocab:
        def __len__(self):
            return CFG["n_token"]

    cfg = NS(MODEL=NS(num_layers=CFG["n_layer"], num_heads=CFG["n_head"], units=CFG["d_model"],
                      inner_size=CFG["d_inner"], dropout=args.dropout, attention_dropout=args.dropout, same_length=False,
                      clamp_len=-1),
             TRAIN=NS(tgt_length=CFG["tgt_len"], mem_length=CFG["mem_len"]))
    torch.manual_seed(1111)
    model = MemTransformerLM(cfg, Vocab())
    with torch.no_grad():      # init recipe of train.py:291-342
        for n, p in model.named_parameters():
            if n.endswith("layer_norm.weight"):
                p.normal_(1.0, 0.01)
            elif n.endswith(".bias") and p.dim() == 1:
                p.zero_()
            else:
                p.normal_(0.0, 0.01)
    model = model.to(dev)
    if args.decode_only:           # development aid: only the BASELINE configs[3] decode arm
        print(json.dumps({"decode": decode_bench(model, dev, peaks())}), flush=True)
        return
    model.train()
    tr = Trainer(model, lr=0.004 / world, warmup_step=100, lr_min=1e-4, clip=1.0, batch_chunk=args.batch_chunk,
                 world=world, comm=comm, global_lr=0.004)
    T, B = CFG["tgt_len"], B_PER_GPU
    K, W = args.steps, args.warmup
    gen = torch.Generator().manual_seed(1111 + 1000 * rank)
    n_batches = 4
    hd, ht = synthetic_tokens(T, B, n_batches, gen)
    hd, ht = hd.pin_memory(), ht.pin_memory()
    dd, dt = hd.to(dev), ht.to(dev)
    reset_h = torch.zeros(B, dtype=torch.bool).pin_memory()
    reset_d = reset_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- warm-up (fills the recurrent memory, compiles nothing: all kernels are prebuilt) ----
    for s in range(W):
        tr.train_step(dd[s % n_batches], dt[s % n_batches], reset_d)
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    lib = nv.lib()
    # ---- device-resident arm ----
    lib.commu_prof_arm(0b0111)
    lib.commu_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for s in range(K):
        loss, _ = tr.train_step(dd[s % n_batches], dt[s % n_batches], reset_d)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    launches = int(lib.commu_launch_count(1))
    import ctypes
    prof = {}
    for cls, name in ((0, "gemm"), (1, "attn_fwd"), (2, "attn_bwd")):
        ms, n = ctypes.c_float(0), ctypes.c_int(0)
        nv.check(lib.commu_prof_read(cls, ctypes.byref(ms), ctypes.byref(n)))
        prof[name] = dict(ms=ms.value, launches=n.value)
    lib.commu_prof_arm(0)
    final_loss = float(loss)

    # ---- end-to-end arm: host inputs in, loss out, every step ----
    barrier()
    e0.record()
    for s in range(K):
        d_ = hd[s % n_batches].to(dev, non_blocking=True)
        t_ = ht[s % n_batches].to(dev, non_blocking=True)
        r_ = reset_h.to(dev, non_blocking=True)
        loss, _ = tr.train_step(d_, t_, r_)
        loss_host = loss.item()
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.window(t_wall0, t_wall1) if sampler else None
    if sampler:
        sampler.stop()

    if rank == 0:
        pk = peaks()
        tokens_step = T * B * world
        value = tokens_step * K / (ms_dev / 1e3)
        e2e = tokens_step * K / (ms_e2e / 1e3)
        f_fwd, f_train, attn_bh = algorithmic_flops_per_token(CFG["n_layer"], CFG["d_model"], CFG["d_inner"],
                                                              T, CFG["mem_len"], B, CFG["n_token"])
        dom = max(("attn_fwd", "attn_bwd", "gemm"), key=lambda k: prof[k]["ms"])
        if dom == "attn_bwd":
            fl = 2 * attn_bh * B * CFG["n_head"]
        elif dom == "attn_fwd":
            fl = attn_bh * B * CFG["n_head"]
        else:
            fl = None
        roof = None
        if fl is not None and prof[dom]["launches"]:
            dur = prof[dom]["ms"] / prof[dom]["launches"] / 1e3
            ach = fl / dur / 1e12
            traf = ncu_traffic(dom) if (args.config == "c2" and B == 16) else None   # captured at the headline shape only
            tens = dict(achieved=round(ach, 2), peak=pk["bf16"], unit="TFLOP/s", frac=round(ach / pk["bf16"], 4),
                        flops_per_launch=fl)
            unit_names = {"attn_bwd": "one layer's attention backward (commu_relattn_bwd) = delta + pass 1 (dS, dK, dV) "
                                      "+ the three band GEMMs (dq_A, dq_C, dR)",
                          "attn_fwd": "one layer's attention forward"}
            if dom == "attn_bwd":
                # the materialised backward is HBM-bound by design: algorithmic HBM traffic = P~ read + dS written once +
                # dS read back twice (by the merged dq_A / dq_C launch, whose second view of the rows is served from L2,
                # and by the dR GEMM) = 8 bytes per causal-visible score element
                elts = B * CFG["n_head"] * (T * CFG["mem_len"] + T * (T + 1) / 2)
                hb = 8.0 * elts
                roof = dict(kernel=dom, bound="hbm", achieved=round(hb / dur / 1e9, 1), peak=pk["hbm"], unit="GB/s",
                            frac=round(hb / dur / 1e9 / pk["hbm"], 4), traffic=traf,
                            peak_source=pk["source"].replace("sustained", "copy bandwidth"),
                            bytes_per_launch=int(hb), bytes_per_unit="8 B per causal-visible score element (bf16 P~ read, bf16 dS written, dS read twice)",
                            avg_launch_ms=round(dur * 1e3, 4), launch_unit=unit_names[dom], tensor_view=tens)
            else:
                roof = dict(kernel=dom, bound="tensor", traffic=traf, peak_source=pk["source"],
                            avg_launch_ms=round(dur * 1e3, 4), launch_unit=unit_names.get(dom, dom), **tens)
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "tokens/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": round(ms_dev / K, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD + " train step (fwd+bwd+clip+Adam), dropout %g / attention_dropout %g "
                                   "(reference default 0.1), batch_chunk %d" % (args.dropout, args.dropout, args.batch_chunk),
                       "global_batch": B * world, "per_gpu_batch": B, "seq_len": T, "mem_len": CFG["mem_len"],
                       "parallelism": "dp%d" % world,
                       "l2": "per-step working set (saved activations + stored attention probabilities: tens of GB per "
                             "GPU) >> 126 MB L2"},
            "e2e": {"value": round(e2e, 1), "unit": "tokens/s", "ms_per_step": round(ms_e2e / K, 3),
                    "h2d_bytes_per_step": int(2 * T * B * 8 + B), "d2h_bytes_per_step": 4},
            "gpu_launches": launches,
            "roofline": roof,
            "step_mfu": {"algorithmic_gflop_per_token": round(f_train / 1e9, 4),
                         "achieved_tflops_per_gpu": round(value / world * f_train / 1e12, 1),
                         "frac_of_bf16_peak": round(value / world * f_train / 1e12 / pk["bf16"], 4)},
            "kernel_time_ms_per_step": {k: round(v["ms"] / K, 3) for k, v in prof.items()},
            "clocks": clocks,
            "final_loss": round(final_loss, 5),
        }
        if world == 1 and not args.no_decode and args.config == "c2":
            try:
                out["decode"] = decode_bench(model, dev, pk)
            except Exception as e:      # the decode arm must never hide the training metric
                out["decode"] = {"error": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline and args.config == "c2":
            out["cpu_baseline"] = cpu_baseline(1, 1, args.dropout)
        if world == 1 and not args.no_reference_gpu and args.config == "c2":
            del tr, model
            torch.cuda.empty_cache()
            try:
                rg = reference_gpu_leg(dev, args.dropout)
                out["reference_gpu"] = rg
                out["vs_reference_gpu"] = {
                    "fp32": round(value / rg["best_fp32_tokens_per_s"], 2) if rg.get("best_fp32_tokens_per_s") else None,
                    "autocast_bf16": round(value / rg["best_autocast_bf16_tokens_per_s"], 2)
                    if rg.get("best_autocast_bf16_tokens_per_s") else None,
                    "e2e_fp32": round(e2e / rg["best_fp32_tokens_per_s"], 2) if rg.get("best_fp32_tokens_per_s") else None,
                    "target": ">= 10x the reference PyTorch-GPU (fp32 eager) tokens/s (BASELINE.json north_star)"}
            except Exception as e:
                out["reference_gpu"] = {"error": repr(e)[:300]}
        print(json.dumps(out), flush=True)
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()


def run_check(args):
    """bench.py --gpus N --check (under torch.distributed.run, N >= 2): on-hardware correctness of the data-parallel
    path (SURVEY.md 8(a) row a14 / section 4 "distributed").
      1. gradient exchange: the flat gradient after the native NCCL exchange (per-layer all-reduces overlapped with the
         backward, and the single all-reduce) == the sum of the ranks' local gradients (gathered with torch.distributed).
      2. N-rank data-parallel training == one GPU on the same GLOBAL batch: per-step loss and the weight update after a
         few optimizer steps (reference semantics: DDP averages the gradients, train.py:155, 467-473)."""
    import torch.distributed as dist
    from types import SimpleNamespace as NS
    from commu.engine.trainer import GradComm, Trainer
    from commu.model.model import MemTransformerLM
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world < 2:
        raise SystemExit("bench.py --check needs >= 2 ranks (python -m torch.distributed.run --nproc-per-node 2 ...)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    comm = GradComm(rank, world, dev)
    L, H, d, Di, T, M, b, V = 3, CFG["n_head"], CFG["d_model"], CFG["d_inner"], 256, 256, 2, CFG["n_token"]

    class Vocab:
        def __len__(self):
            return V
    cfg = NS(MODEL=NS(num_layers=L, num_heads=H, units=d, inner_size=Di, dropout=0.0, attention_dropout=0.0,
                      same_length=False, clamp_len=-1), TRAIN=NS(tgt_length=T, mem_length=M))

    def fresh_model():
        torch.manual_seed(1234)                       # identical weights on every rank
        m = MemTransformerLM(cfg, Vocab())
        with torch.no_grad():
            for n, p in m.named_parameters():
                if n.endswith("layer_norm.weight"):
                    p.normal_(1.0, 0.02)
                elif n.endswith(".bias") and p.dim() == 1:
                    p.zero_()
                else:
                    p.normal_(0.0, 0.02)
        return m.to(dev).train()
    gen = torch.Generator().manual_seed(99)
    steps = 3
    tok = torch.randint(2, 560, (steps + 1, T + 1, b * world), generator=gen)
    glob = [(tok[s, :-1].contiguous().to(dev), tok[s, 1:].contiguous().to(dev)) for s in range(steps + 1)]
    mine = slice(rank * b, (rank + 1) * b)
    res = {"check": "ok", "world": world, "shape": "L%d d%d H%d Di%d T%d M%d, %d columns per rank" % (L, d, H, Di, T, M, b)}
    # ---- 1. gradient exchange ----
    tr = Trainer(fresh_model(), lr=0.004, warmup_step=0, lr_min=1e-4, clip=1.0, world=world, comm=comm)
    tr.accumulate_gradients(glob[0][0][:, mine].contiguous(), glob[0][1][:, mine].contiguous(), None)   # fills the memory
    tr.apply_update()
    saved_mems = list(tr.mems)
    d1, t1 = glob[1][0][:, mine].contiguous(), glob[1][1][:, mine].contiguous()
    tr.accumulate_gradients(d1, t1, None, exchange=False)
    g_local = tr.flat_g.clone()
    parts = [torch.empty_like(g_local) for _ in range(world)]
    dist.all_gather(parts, g_local)
    g_sum = torch.stack(parts).double().sum(0)
    scale = float(g_sum.abs().max())
    for mode, overlap in (("overlapped per-layer all-reduces", True), ("single all-reduce", False)):
        tr.mems = list(saved_mems)
        tr.overlap = overlap
        tr.accumulate_gradients(d1, t1, None, exchange=True)
        torch.cuda.synchronize()
        err = float((tr.flat_g.double() - g_sum).abs().max())
        # (the backward accumulates weight gradients with fp32 atomics: two runs of the SAME rank differ at this level)
        tr.mems = list(saved_mems)
        tr.accumulate_gradients(d1, t1, None, exchange=False)
        noise = float((tr.flat_g - g_local).abs().max())
        res["exchange: " + mode] = {"max_abs_err": err, "grad_max": scale, "run_to_run_noise_one_rank": noise}
        if not err <= 1e-4 * scale + 4 * world * noise:
            res["check"] = "FAILED: " + mode
    tr.overlap = True
    # ---- 2. data parallel == one GPU at the same global batch ----
    m_dp = fresh_model()
    p0 = {n: p.detach().clone() for n, p in m_dp.named_parameters()}
    tr_dp = Trainer(m_dp, lr=0.004, warmup_step=0, lr_min=1e-4, clip=1.0, world=world, comm=comm)
    dp_losses = []
    for s in range(steps):
        loss, _ = tr_dp.train_step(glob[s][0][:, mine].contiguous(), glob[s][1][:, mine].contiguous(), None)
        l = loss.detach().clone()
        dist.all_reduce(l)
        dp_losses.append(float(l) / world)
    if rank == 0:
        m_1 = fresh_model()
        tr_1 = Trainer(m_1, lr=0.004, warmup_step=0, lr_min=1e-4, clip=1.0)
        one_losses = [float(tr_1.train_step(glob[s][0], glob[s][1], None)[0]) for s in range(steps)]
        num = den = 0.0
        for (n, a), (_, c) in zip(m_dp.named_parameters(), m_1.named_parameters()):
            num += float(((a - c).double() ** 2).sum())
            den += float(((c - p0[n]).double() ** 2).sum())
        upd_err = (num / max(den, 1e-30)) ** 0.5
        loss_err = max(abs(x - y) / y for x, y in zip(dp_losses, one_losses))
        res["data_parallel_vs_one_gpu"] = {"dp_losses": dp_losses, "one_gpu_losses": one_losses,
                                           "max_rel_loss_err": loss_err, "rel_err_of_weight_update": upd_err}
        if not (loss_err < 1e-3 and upd_err < 0.05):
            res["check"] = "FAILED: data parallel vs one GPU"
        print(json.dumps(res), flush=True)
    ok = torch.tensor([1 if res["check"] == "ok" else 0], device=dev)
    dist.broadcast(ok, 0)
    comm.close()
    dist.destroy_process_group()
    if int(ok) != 1:
        raise SystemExit(1)


def decode_cpu_baseline(n_tokens=16, mem_len=2048):
    """The reference decode path (forward_generate over [memory ; token] per generated token,
    midi_inferrer.py:199-207) on the host cores through the oracle port: batch 1, memory pre-filled, greedy."""
    from oracle import transfoxl_oracle as orc
    nthreads = pick_cpu_threads()
    cfg = orc.make_cfg(CFG["n_layer"], CFG["n_head"], CFG["d_model"], CFG["d_inner"], 1, mem_len, True, -1, CFG["n_token"])
    P = orc.init_params(cfg, seed=1111, std=0.01)
    gen = torch.Generator().manual_seed(3)
    ctx = torch.randint(2, 560, (mem_len + 1, 1), generator=gen)
    with torch.no_grad():
        _, mems = orc.forward_generate(cfg, P, ctx[:-1], None)
        cur = ctx[-1:]
        t0 = time.time()
        for _ in range(n_tokens):
            lg, mems = orc.forward_generate(cfg, P, cur, mems)
            cur = (1 + lg[-1, :, 1:].argmax(-1))[None]
        dt = time.time() - t0
    return {"value": round(n_tokens / dt, 2), "unit": "tokens/s/seq", "cores": nthreads, "kind": "port",
            "os_cpu_count": os.cpu_count(),
            "sample": "%d greedy tokens, batch 1, memory %d pre-filled, %.1f s" % (n_tokens, mem_len, dt)}


def decode_bench(model, dev, pk, n_new=512, batch=64, mem_len=2048):
    """BASELINE configs[3]: batch 64, 12L d512, mem_len 2048 (pre-filled), top-p 0.9 / T 0.95, 512 new tokens.
    The headline is the fp32 engine - the one whose greedy tokens are identical to the fp32 reference
    (tests/test_decode_gpu.py at this shape); the bf16 engine (half the streamed bytes, NOT token-exact) is reported
    beside it as the throughput mode.  `e2e`: the same loop with every step's token ids copied to pinned host memory;
    `cpu_baseline`: the reference decode path on the host cores (oracle port, batch 1, 16 tokens)."""
    res = _decode_one(model, dev, pk, n_new, batch, mem_len, "fp32")
    try:
        res["bf16_throughput_mode"] = _decode_one(model, dev, pk, n_new, batch, mem_len, "bf16")
    except Exception as e:
        res["bf16_throughput_mode"] = {"error": repr(e)[:200]}
    try:
        res["cpu_baseline"] = decode_cpu_baseline(16, mem_len)
    except Exception as e:
        res["cpu_baseline"] = {"error": repr(e)[:200]}
    return res


def _decode_one(model, dev, pk, n_new, batch, mem_len, precision):
    import ctypes
    from commu import _native as nv
    from commu.engine.decode import DecodeEngine, DecodeState
    lib = nv.lib()
    model.eval()
    eng = DecodeEngine(model, batch=batch, mem_len=mem_len, same_length=True, precision=precision)
    gen = torch.Generator().manual_seed(5)
    state = DecodeState()
    fill = torch.randint(2, 560, (mem_len + 8, batch), generator=gen).to(dev)
    if os.environ.get("COMMU_BENCH_FAST_PREFILL") == "1":   # profiling aid (ncu): pretend the ring is already full
        state = DecodeState(mem_len, mem_len - 1)
        fill = fill[:4]
    for t in range(fill.shape[0]):                      # pre-fill the ring cache
        logits, state = eng.step(fill[t].contiguous(), state)
    cur, _ = eng.sample(logits, 0.95, 0, 0.9, None, 1, 0)
    torch.cuda.synchronize()
    # the timed loop replays ONE captured CUDA graph per token (embed, 12 layers, logits, sampler);
    # ring slot / visible count / sampler counter are device-resident (commu_decode_advance)
    dstate = torch.tensor([state.slot, 0, state.count, 0], dtype=torch.int32, device=dev)

    def one_step():
        nv.call("commu_decode_advance", dstate, eng.C, eng.mem_len, 0)
        eng._step_kernels(cur, 0, 1, dstate)
        nv.call("commu_sample", eng.ws["logits"], eng.V, batch, eng.V, 0.95, 0, 0.9, None, 1, 0, cur, None, eng.V, dstate)

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        one_step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        one_step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    lib.commu_prof_arm(0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(n_new):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    # end to end: every step's token ids leave the device (pinned host buffer, one async copy per step)
    host_tokens = torch.empty(n_new, batch, dtype=torch.int64).pin_memory()
    e0.record()
    for t in range(n_new):
        graph.replay()
        host_tokens[t].copy_(cur, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    assert int(host_tokens.min()) >= 1 and int(host_tokens.max()) < eng.V
    # duration of the decode-attention kernel: the 12 layers' launches captured alone in a CUDA graph (each layer streams
    # its own 268 MB cache, far above the 126 MB L2) and replayed back to back with events around the replays, so the
    # per-launch figure carries the same launch gaps as the real step and no host-side launch latency
    lib.commu_prof_arm(0)
    pms, pn = ctypes.c_float(0), ctypes.c_int(0)
    if getattr(eng, "fused", False):
        def attn_chain():
            for l in range(eng.L):
                eng._attn_fused(l, 0, 1, dstate)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            attn_chain()
        torch.cuda.current_stream().wait_stream(side)
        g2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g2):
            attn_chain()
        for _ in range(3):
            g2.replay()
        torch.cuda.synchronize()
        reps = 20
        e0.record()
        for _ in range(reps):
            g2.replay()
        e1.record()
        torch.cuda.synchronize()
        pms.value, pn.value = e0.elapsed_time(e1), reps * eng.L
    else:       # kernel-per-op engines: events around every launch of eager steps
        lib.commu_prof_arm(0b1000)
        for t in range(8):
            one_step()
        torch.cuda.synchronize()
        nv.check(lib.commu_prof_read(3, ctypes.byref(pms), ctypes.byref(pn)))
        lib.commu_prof_arm(0)
    esz = 2 if precision == "bf16" else 4
    L, H, d = CFG["n_layer"], CFG["n_head"], CFG["d_model"]
    attn_bytes = batch * H * mem_len * 64 * esz * 2 + H * mem_len * 64 * esz     # K + V per sequence, R once (8d)
    step_bytes = L * attn_bytes + 41.3e6 * esz
    res = {"tokens_per_s_per_seq": round(n_new / (ms / 1e3), 1), "batch": batch, "new_tokens": n_new,
           "e2e": {"value": round(n_new / (ms_e2e / 1e3), 1), "unit": "tokens/s/seq", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": batch * 8},
           "token_exact_vs_fp32_reference": precision == "fp32",
           "ms_per_token_step": round(ms / n_new, 4), "precision": precision, "sampler": "top_p=0.9 T=0.95",
           "aggregate_tokens_per_s": round(n_new * batch / (ms / 1e3), 1),
           "step_hbm_frac": round(step_bytes / (ms / n_new / 1e3) / 1e9 / pk["hbm"], 4)}
    if pn.value:
        dur = pms.value / pn.value / 1e3
        res["roofline"] = {"kernel": "decode_attn", "bound": "hbm", "achieved": round(attn_bytes / dur / 1e9, 1),
                           "peak": pk["hbm"], "unit": "GB/s", "frac": round(attn_bytes / dur / 1e9 / pk["hbm"], 4),
                           "traffic": ncu_traffic("decode_attn"), "avg_launch_ms": round(dur * 1e3, 4)}
    return res


def pick_cpu_threads():
    """Thread count of the CPU arms, PINNED (no timing-based calibration: that made the count, and with it the
    baseline, vary from run to run): the CPU share the container is actually given - the cgroup quota if one is set,
    else the affinity mask - capped at 32 (the box reports far more logical CPUs than the container may use; a
    128-thread run was measured ~500x slower than a 16-thread run on the same host).  COMMU_CPU_THREADS overrides."""
    n_cpu = os.cpu_count() or 1
    try:
        n_cpu = min(n_cpu, len(os.sched_getaffinity(0)))
    except (AttributeError, OSError):
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n_cpu = min(n_cpu, max(1, int(math.ceil(int(quota) / int(period)))))
    except (OSError, ValueError):
        pass
    n = int(os.environ.get("COMMU_CPU_THREADS", "0")) or min(n_cpu, 32 if n_cpu <= 32 else 16)
    torch.set_num_threads(n)
    return n


def cpu_baseline(steps, warmup, p_drop=0.1):
    """Reference algorithm on the host cores: oracle port (torch CPU fp32, auto-calibrated thread count),
    bounded sample = B=1 sequence of 2048 tokens per step with the memory carried over."""
    from oracle import transfoxl_oracle as orc
    pick_cpu_threads()
    cfg = orc.make_cfg(CFG["n_layer"], CFG["n_head"], CFG["d_model"], CFG["d_inner"], CFG["tgt_len"],
                       CFG["mem_len"], False, -1, CFG["n_token"])
    P = orc.init_params(cfg, seed=1111, std=0.01)
    opt = orc.AdamState(P)
    gen = torch.Generator().manual_seed(7)
    mems = [None]
    T = CFG["tgt_len"]
    times = []
    drop = None
    if p_drop > 0:                                     # the reference's nn.Dropout modules (train mode)
        def drop(site, layer, t):
            return torch.nn.functional.dropout(t, p_drop, True)
    for s in range(warmup + steps):
        tok = torch.randint(2, 560, (T + 1, 1), generator=gen)
        t0 = time.time()
        _, _, mems, _ = orc.train_step(cfg, P, opt, [(tok[:-1], tok[1:], torch.zeros(1, dtype=torch.bool))],
                                       mems, lr=1e-4, drop=drop)
        if s >= warmup:
            times.append(time.time() - t0)
    tot = sum(times)
    return {"value": round(T * len(times) / tot, 2), "unit": "tokens/s", "cores": torch.get_num_threads(),
            "os_cpu_count": os.cpu_count(), "kind": "port", "sample": "%d step(s) of B=1 x T=2048 (M=2048) fwd+bwd+clip+Adam, dropout %g, %.1f s" %
                                      (len(times), p_drop, tot)}


def reference_gpu_leg(dev, p_drop, steps=3, warmup=2):
    """The number the north-star ratio is about (SURVEY.md 8(d) "Reference GPU baseline"): the reference algorithm
    (oracle port of commu/model/model.py:540-693, plain eager PyTorch) with the restated train loop of
    train.py:133-169 (pad-masked mean loss, backward, clip_grad_norm_(1.0), torch.optim.Adam) on THIS B200, same
    model / T / M / dropout, synthetic tokens: eager fp32 (what the reference runs: it never enables autocast or
    TF32) and, labelled, under torch.autocast(bf16) as a stronger baseline.  B = 4 (like-for-like micro-batch of
    SURVEY 8d) and the largest of 8 / 16 that fits the 180 GB.  The oracle is only the thing timed here, never part
    of the product path."""
    from oracle import transfoxl_oracle as orc
    cfg = orc.make_cfg(CFG["n_layer"], CFG["n_head"], CFG["d_model"], CFG["d_inner"], CFG["tgt_len"],
                       CFG["mem_len"], False, -1, CFG["n_token"])
    T = CFG["tgt_len"]
    res = {}

    def drop(site, layer, t):
        return torch.nn.functional.dropout(t, p_drop, True)

    def one(B, autocast):
        P0 = orc.init_params(cfg, seed=1111, std=0.01)
        P = {k: torch.nn.Parameter(v.to(dev)) for k, v in P0.items()}
        opt = torch.optim.Adam(list(P.values()), lr=1e-4)
        gen = torch.Generator(device=dev).manual_seed(7)
        mems = None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for s in range(warmup + steps):
            if s == warmup:
                torch.cuda.synchronize()
                e0.record()
            tok = torch.randint(2, 560, (T + 1, B), generator=gen, device=dev)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                nll, mems = orc.forward_loss(cfg, P, tok[:-1], tok[1:], None, mems, drop=drop if p_drop > 0 else None)
                loss = nll[tok[1:] != 0].float().mean()
            loss.backward()
            torch.nn.utils.clip_grad_norm_(list(P.values()), 1.0)
            opt.step()
            opt.zero_grad(set_to_none=True)
            loss.item()                               # the reference logs the loss every step (train.py:157)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"tokens_per_s": round(T * B / (ms / 1e3), 1), "ms_per_step": round(ms, 2), "batch": B,
                "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)}

    for B in (4, 16, 8):
        if B == 8 and "fp32_B16" in res and "error" not in res["fp32_B16"]:
            continue
        for autocast in (False, True):
            key = ("autocast_bf16" if autocast else "fp32") + "_B%d" % B
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
            try:
                res[key] = one(B, autocast)
            except torch.cuda.OutOfMemoryError:
                res[key] = {"error": "out of memory at B=%d" % B}
            except Exception as e:                     # never hide the training metric behind the baseline
                res[key] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
    ok = lambda pre: [v["tokens_per_s"] for k, v in res.items() if k.startswith(pre) and "tokens_per_s" in v]
    res["best_fp32_tokens_per_s"] = max(ok("fp32"), default=None)
    res["best_autocast_bf16_tokens_per_s"] = max(ok("autocast"), default=None)
    res["what"] = ("oracle port of the reference model + restated train.py loop, eager PyTorch %s on this GPU, "
                   "dropout %g, %d timed steps after %d warm-up" % (torch.__version__, p_drop, steps, warmup))
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    t0 = time.time()
    cb = cpu_baseline(K, W, args.dropout)
    out = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "tokens/s",
           "n_gpus": args.gpus, "steps": K, "warmup": W,
           "ms_per_step": round(CFG["tgt_len"] / cb["value"] * 1e3, 1), "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "BASELINE configs[1] on CPU (oracle port of the reference algorithm): 12L d512 "
                                  "H8 Di2048 T=2048 M=2048, dropout %g, bounded sample B=1 per step" % args.dropout, "global_batch": 1,
                      "seq_len": CFG["tgt_len"], "parallelism": "cpu"},
           "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0, "wall_s": round(time.time() - t0, 1)}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="c2 = BASELINE configs[1] (headline, default); c5 = configs[4], 24L d1024 T=M=4096, B=2 per GPU")
    ap.add_argument("--check", action="store_true",
                    help="multi-GPU correctness instead of timing: all-reduced gradient == sum of the ranks' gradients, "
                         "data-parallel loss / weights == one GPU at the same global batch")
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--dropout", type=float, default=0.1,
                    help="MODEL.dropout = MODEL.attention_dropout (reference default 0.1, config_helper.py:11-12)")
    ap.add_argument("--batch-chunk", type=int, default=1,
                    help="TRAIN.batch_chunk micro-batches per optimizer step (reference default 4, train.py:123,136-155)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-decode", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--decode-only", action="store_true")
    args = ap.parse_args()
    select_config(args.config)
    if args.check:
        return run_check(args)
    if args.warmup < 3 and args.impl == "native" and os.environ.get("COMMU_BENCH_PROFILE") != "1":
        args.warmup = 3          # (COMMU_BENCH_PROFILE=1: launch-list runs under ncu, whose numbers are never reported)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_native(args)


if __name__ == "__main__":
    # stdout carries the JSON line(s) only: native libraries that write to fd 1 (NCCL's version banner under torchrun)
    # are pointed at stderr, Python's own stdout keeps the original descriptor
    sys.stdout.flush()
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, "w")
    main()
