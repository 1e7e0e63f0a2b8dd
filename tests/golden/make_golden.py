"""Generates tests/golden/*.npz by importing the UNMODIFIED reference (POZAlabs/ComMU-code) from
/root/reference in the build container.  The reference cannot travel to the GPU box, so the
outputs are frozen here; this script is committed so the fixtures are reproducible.

    python tests/golden/make_golden.py            # rewrites every fixture

Third-party modules the reference imports at module scope but that are absent from this image
(yacs, miditoolkit, parmap, pretty_midi) are stubbed in sys.modules ONLY to make the import of
commu.midi_generator.midi_inferrer succeed; none of them is on the arithmetic path.
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = os.environ.get("COMMU_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _stub_modules():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    class _CfgNode(dict):
        pass

    mod("yacs")
    mod("yacs.config", CfgNode=_CfgNode)
    sys.modules["yacs"].config = sys.modules["yacs.config"]
    mod("miditoolkit", MidiFile=object, Instrument=object, TempoChange=object, Note=object,
        Marker=object, TimeSignature=object)
    mod("miditoolkit.midi")
    mod("miditoolkit.midi.parser", MidiFile=object)
    mod("miditoolkit.midi.containers", Marker=object, TimeSignature=object, TempoChange=object,
        Instrument=object, Note=object)
    mod("parmap")
    mod("pretty_midi")
    mod("splitfolders")


def ref_cfg(n_layer, n_head, d_model, d_inner, tgt_len, mem_len, same_length, clamp_len):
    return SimpleNamespace(
        MODEL=SimpleNamespace(num_layers=n_layer, num_heads=n_head, units=d_model,
                              inner_size=d_inner, dropout=0.0, attention_dropout=0.0,
                              same_length=same_length, clamp_len=clamp_len),
        TRAIN=SimpleNamespace(tgt_length=tgt_len, mem_length=mem_len))


class Vocab:
    def __init__(self, n):
        self.n = n

    def __len__(self):
        return self.n


def build_ref_model(cfgd, n_token, seed, std):
    """Reference model + the init recipe of train.py:291-342 (restated; train.py itself is not
    importable because it initialises NCCL at import)."""
    from commu.model.model import MemTransformerLM
    torch.manual_seed(seed)
    model = MemTransformerLM(ref_cfg(**cfgd), Vocab(n_token))
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("layer_norm.weight"):
                p.normal_(1.0, std)
            elif name.endswith(".bias") and p.dim() == 1:
                p.zero_()
            else:
                p.normal_(0.0, std)
    return model


def state_np(model):
    return {"param/" + k: v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()
            if k != "crit.out_layers.0.weight"}


def gen_forward_case(name, cfgd, n_token, B, n_seg, seed, std, reset_plan):
    model = build_ref_model(cfgd, n_token, seed, std)
    model.eval()
    rng = np.random.RandomState(seed)
    T = cfgd["tgt_len"]
    out = dict(state_np(model))
    out["cfg"] = np.array([cfgd["n_layer"], cfgd["n_head"], cfgd["d_model"], cfgd["d_inner"],
                           cfgd["tgt_len"], cfgd["mem_len"], int(cfgd["same_length"]),
                           cfgd["clamp_len"], n_token], dtype=np.int64)
    mems = None
    model.zero_grad()
    for s in range(n_seg):
        data = torch.from_numpy(rng.randint(1, n_token, size=(T, B))).long()
        target = torch.from_numpy(rng.randint(0, n_token, size=(T, B))).long()
        reset = torch.tensor(reset_plan[s], dtype=torch.bool) if reset_plan else None
        loss, mems = model(data, target, reset, mems)
        (loss.mean()).backward()
        out["seg%d/data" % s] = data.numpy()
        out["seg%d/target" % s] = target.numpy()
        out["seg%d/reset" % s] = (reset.numpy() if reset is not None else np.zeros(B, bool))
        out["seg%d/loss" % s] = loss.detach().numpy()
        out["seg%d/mems" % s] = mems.detach().numpy()
    for k, p in model.named_parameters():
        if k == "crit.out_layers.0.weight":
            continue
        out["grad/" + k] = p.grad.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "M_final", mems.shape[1])


def gen_decode_case(name, cfgd, n_token, B, n_ctx, n_new, seed, std):
    model = build_ref_model(cfgd, n_token, seed, std)
    model.eval()
    model.reset_length(1, cfgd["mem_len"])
    rng = np.random.RandomState(seed + 1)
    out = dict(state_np(model))
    out["cfg"] = np.array([cfgd["n_layer"], cfgd["n_head"], cfgd["d_model"], cfgd["d_inner"],
                           1, cfgd["mem_len"], int(cfgd["same_length"]), cfgd["clamp_len"],
                           n_token], dtype=np.int64)
    ctx = torch.from_numpy(rng.randint(1, n_token, size=(n_ctx, B))).long()
    toks, logits_all = [], []
    with torch.no_grad():
        _, mems = model.forward_generate(ctx[:-1], None)          # context prefill (multi-token)
        cur = ctx[-1:]
        for _ in range(n_new):
            lg, mems = model.forward_generate(cur, mems)
            nxt = 1 + lg[-1, :, 1:].argmax(dim=-1)               # greedy, token 0 never sampled
            toks.append(nxt.numpy())
            logits_all.append(lg[-1].numpy())
            cur = nxt[None, :]
    out["ctx"] = ctx.numpy()
    out["tokens"] = np.stack(toks)              # [n_new, B]
    out["logits"] = np.stack(logits_all)        # [n_new, B, V]
    out["final_mems"] = mems.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    srt = np.sort(out["logits"][:, :, 1:], axis=-1)
    print("wrote", name, "min greedy margin", float((srt[..., -1] - srt[..., -2]).min()))


def gen_sampler_case(name, seed):
    _stub_modules()
    from commu.midi_generator.midi_inferrer import InferenceTask
    rng = np.random.RandomState(seed)
    out = {}
    for ci, (temp, top_k, wrong) in enumerate([(0.95, 32, []), (0.95, 32, [5, 300]), (1.3, 8, [17]),
                                               (0.0, 32, [])]):
        task = InferenceTask(torch.device("cpu"))
        task.input_data = SimpleNamespace(temperature=temp, top_k=top_k)
        full = torch.from_numpy((rng.randn(729) * 3.0).astype(np.float32))
        logits = full[1:].clone()                 # what calc_logits_and_mems returns (:206)
        probs = task.calc_probs(logits)
        probs = task.apply_sampling(probs, wrong)
        out["case%d/logits_full" % ci] = full.numpy()
        out["case%d/params" % ci] = np.array([temp, top_k], dtype=np.float64)
        out["case%d/wrong" % ci] = np.array(wrong, dtype=np.int64)
        out["case%d/probs" % ci] = probs.numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


def gen_train_case(name, cfgd, n_token, B, chunks, n_steps, seed, std, lr, warmup, lr_min=1e-4):
    """Restates train.py:133-169 + 441-461 around the reference model with torch.optim.Adam."""
    model = build_ref_model(cfgd, n_token, seed, std)
    model.train()
    rng = np.random.RandomState(seed + 7)
    T = cfgd["tgt_len"]
    out = dict(state_np(model))
    out["cfg"] = np.array([cfgd["n_layer"], cfgd["n_head"], cfgd["d_model"], cfgd["d_inner"],
                           cfgd["tgt_len"], cfgd["mem_len"], int(cfgd["same_length"]),
                           cfgd["clamp_len"], n_token], dtype=np.int64)
    out["hyper"] = np.array([lr, warmup, lr_min, 1.0, chunks], dtype=np.float64)
    opt = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=0.0)

    def lam(step):
        if step == 0 and warmup == 0:
            return 1.0
        return max((warmup ** 0.5) / (step ** 0.5), lr_min / lr) if step > warmup else step / warmup

    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lam)
    mems = [None] * chunks
    losses, gnorms, lrs = [], [], []
    for step in range(n_steps):
        data = torch.from_numpy(rng.randint(1, n_token, size=(T, B))).long()
        target = torch.from_numpy(rng.randint(0, n_token, size=(T, B))).long()
        reset = torch.from_numpy(rng.rand(B) < 0.2)
        out["step%d/data" % step] = data.numpy()
        out["step%d/target" % step] = target.numpy()
        out["step%d/reset" % step] = reset.numpy()
        model.zero_grad()
        tot = 0.0
        dc, tc, rc = torch.chunk(data, chunks, 1), torch.chunk(target, chunks, 1), torch.chunk(reset, chunks, 0)
        for i in range(chunks):
            loss, mems[i] = model(dc[i].contiguous(), tc[i].contiguous(), rc[i].contiguous(), mems[i])
            loss = loss[tc[i] != 0].float().mean() / chunks
            tot += loss.item()
            loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        opt.zero_grad()
        sched.step()
        losses.append(tot)
        gnorms.append(float(gn))
    out["losses"] = np.array(losses)
    out["gnorms"] = np.array(gnorms)
    out["lrs"] = np.array(lrs)
    for k, v in model.state_dict().items():
        if k != "crit.out_layers.0.weight":
            out["final/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name, "losses", losses)


def gen_dataset_case(name, seed):
    """Batches produced by the reference ComMUDataset on a synthetic corpus (file format of
    SURVEY.md 8d, written by the new repo's writer; the corpus itself is stored in the fixture)."""
    import tempfile
    _stub_modules()
    from commu.model.dataset import ComMUDataset
    rng = np.random.RandomState(seed)
    d = tempfile.mkdtemp()
    out = {}
    for tag, n in (("train", 23), ("val", 7)):
        inp = np.empty(n, dtype=object)
        tgt = np.empty(n, dtype=object)
        for i in range(n):
            ln = int(rng.randint(20, 90))
            inp[i] = rng.randint(560, 729, size=11).astype(np.int64)
            ev = rng.randint(2, 560, size=ln).astype(np.int16)
            ev[-1] = 1
            tgt[i] = ev
            out["corpus/%s/%d/input" % (tag, i)] = inp[i]
            out["corpus/%s/%d/target" % (tag, i)] = tgt[i]
        np.save(os.path.join(d, "input_%s.npy" % tag), inp, allow_pickle=True)
        np.save(os.path.join(d, "target_%s.npy" % tag), tgt, allow_pickle=True)
    ds = ComMUDataset(d, None)
    it = ds.get_iterator(5, 16, "cpu", "train", True, seed=1111)()
    for b in range(40):                      # > one epoch (23 samples x ~55 tokens / (5 x 16))
        data, target, reset, ntok = next(it)
        out["train/%d/data" % b] = data.numpy().copy()
        out["train/%d/target" % b] = target.numpy().copy()
        out["train/%d/reset" % b] = reset.numpy().copy()
        out["train/%d/ntok" % b] = np.array(ntok)
    for rank in range(2):
        for b, (data, target, first, ntok) in enumerate(ds.eval_iterator(3, 16, "cpu", "valid", rank, 2)()):
            out["eval%d/%d/data" % (rank, b)] = data.numpy().copy()
            out["eval%d/%d/target" % (rank, b)] = target.numpy().copy()
            out["eval%d/%d/first" % (rank, b)] = np.array(first)
            out["eval%d/%d/ntok" % (rank, b)] = np.array(ntok)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


TEACHER_SCENARIOS = {
    # name: (num_measures, chord tokens, chord positions, scripted "model wishes")
    "one_per_bar": (4, [200, 210, 220, 230], [432, 432, 432, 432],
                    [2, 440, 150, 60, 310, 2, 450, 140, 62, 320, 250, 2, 433, 160, 64, 330, 1, 2, 470, 135, 66, 340, 2, 1]),
    "inter_chords": (4, [201, 211, 221, 231, 241, 251], [432, 496, 432, 432, 480, 432],
                     [2, 440, 150, 60, 310, 500, 150, 61, 311, 2, 450, 140, 62, 320, 2, 260, 470, 160, 64, 330, 490, 2, 433, 135, 66, 340, 1, 2, 1]),
    "incomplete": (5, [202, 212, 222, 232], [432, 432, 432, 432],
                   [2, 436, 150, 60, 310, 2, 440, 150, 60, 310, 2, 450, 140, 62, 320, 2, 433, 160, 64, 330, 2, 470, 135, 66, 340, 2, 1]),
}


def fake_model_patch(task, script):
    """Deterministic stand-in for the network + multinomial draw, shared by the golden generator and the
    CPU test: logits favour the next scripted token (plus seeded noise); the draw is the arg-max."""
    state = {"calls": 0, "ptr": 0}

    def calc_logits_and_mems(seq, mems):
        state["calls"] += 1
        g = torch.Generator().manual_seed(1000 + state["calls"])
        base = torch.randn(728, generator=g)
        want = script[state["ptr"]] if state["ptr"] < len(script) else 1
        base[want - 1] += 20.0
        return base, (mems or 0) + 1

    def infer_token(probs):
        state["ptr"] += 1
        return int(torch.argmax(probs))

    task.calc_logits_and_mems = calc_logits_and_mems
    task.infer_token = infer_token
    return state


def gen_teacher_case(name):
    _stub_modules()
    from commu.midi_generator.midi_inferrer import InferenceTask
    out = {}
    for sc, (nm, ctok, cpos, script) in TEACHER_SCENARIOS.items():
        task = InferenceTask(torch.device("cpu"))
        task.input_data = SimpleNamespace(num_measures=nm, temperature=0.95, top_k=32, num_generate=1,
                                          chord_token_components={"chord_token": list(ctok), "chord_position": list(cpos)})
        task.inference_cfg = SimpleNamespace(GENERATION=SimpleNamespace(generation_length=200))
        fake_model_patch(task, script)
        seq = task.generate_sequence([0, 574, 623, 627, 635, 639, 642, 651, 684, 694, 720, 727], 0)
        out[sc] = np.array(seq if seq is not None else [-1], dtype=np.int64)
        print(" teacher", sc, "->", None if seq is None else len(seq))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


GENERATE_META = [574, 623, 627, 635, 639, 642, 651, 684, 694, 720, 727]
GENERATE_SCENARIOS = {
    # name: (weight seed, weight std, logit bias of BAR, num_measures, chord tokens, chord positions, generation_length)
    "a": (1, 0.35, 3.0, 4, [200, 210, 220, 230, 240], [432, 496, 432, 432, 480], 200),
    "b": (4, 0.35, 3.0, 4, [200, 210, 220, 230, 240], [432, 496, 432, 432, 480], 200),
    "c": (1, 0.5, 4.0, 4, [200, 210, 220, 230, 240], [432, 496, 432, 432, 480], 200),     # 67 tokens > mem_len 48
}
GENERATE_CFG = dict(n_layer=2, n_head=4, d_model=64, d_inner=128, tgt_len=1, mem_len=48, same_length=True, clamp_len=-1)


def generate_logit_bias(bar):
    """Output-bias pattern that makes a random tiny model emit bars, positions and an end-of-sequence now and then,
    so that the greedy run walks through the teacher-forcing branches."""
    bias = {2: bar, 1: bar - 1.5}
    bias.update({t: 1.5 for t in range(433, 560, 5)})
    return bias


def gen_generate_case(name):
    """REAL tiny reference model + the reference's own InferenceTask.generate_sequence at temperature 0 (chord teacher
    forcing: first position, one-chord / inter-chord cases, skipped chord position, end of sequence).  The raw
    sequence is recorded before validate_teacher_forced_sequence judges it (a random model rarely produces the
    right number of bars), so the test compares every generated and forced token."""
    _stub_modules()
    from commu.midi_generator.midi_inferrer import InferenceTask, TeacherForceTask
    out = {"meta": np.array(GENERATE_META, dtype=np.int64)}
    for sc, (seed, std, bar, nm, ctok, cpos, gen_len) in GENERATE_SCENARIOS.items():
        model = build_ref_model(GENERATE_CFG, 729, seed, std)
        with torch.no_grad():
            for tok, val in generate_logit_bias(bar).items():
                model.crit.out_layers[0].bias[tok] = val
        model.eval()
        model.reset_length(1, GENERATE_CFG["mem_len"])
        task = InferenceTask(torch.device("cpu"))
        task(model, SimpleNamespace(num_measures=nm, temperature=0.0, top_k=32, num_generate=1,
                                    chord_token_components={"chord_token": list(ctok), "chord_position": list(cpos)}),
             SimpleNamespace(GENERATION=SimpleNamespace(generation_length=gen_len)))
        keep = {}
        orig = TeacherForceTask.validate_teacher_forced_sequence

        def record(self, seq, keep=keep, orig=orig):
            keep["seq"] = list(seq)
            return orig(self, seq)
        TeacherForceTask.validate_teacher_forced_sequence = record
        try:
            with torch.no_grad():
                seq, mems = task.init_seq_and_mems(list(GENERATE_META), len(GENERATE_META))
                verdict = task.generate_sequence(seq, mems)
        finally:
            TeacherForceTask.validate_teacher_forced_sequence = orig
        for k, v in state_np(model).items():
            out[sc + "/" + k] = v
        out[sc + "/raw"] = np.array(keep["seq"], dtype=np.int64)
        out[sc + "/valid"] = np.array([verdict is not None])
        print(" generate", sc, "->", len(keep["seq"]), "tokens,", "valid" if verdict is not None else "rejected by validation")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("wrote", name)


# ---------------------------------------------------------------------------------------------------
# A checkpoint WRITTEN BY THE REFERENCE at its code-default model shape (6 layers, 10 heads, d_model 500,
# d_inner 1000, Dh 50; config_helper.py:7-10) and resumed by the reference.  The file is ~170 MB (weights + Adam
# moments), far too large for a fixture, so its CONTENT is a pure function of the key names (seeded per key, see
# keyed_tensor) and only the structure written by torch.save, the synthetic tokens' seed and the reference's
# continued losses are stored; tests/test_checkpoint_gpu.py rebuilds the identical file from that.
# ---------------------------------------------------------------------------------------------------
CKPT_CFG = dict(n_layer=6, n_head=10, d_model=500, d_inner=1000, tgt_len=128, mem_len=1024, same_length=False,
                clamp_len=-1)
CKPT_HYPER = dict(lr=0.004, warmup=100, lr_min=1e-4, chunks=2, B=4, train_step=137, n_steps=3, data_seed=77)


def keyed_tensor(key, shape, kind):
    """Deterministic tensor content from the key name alone (same torch CPU generator here and in the test)."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(key.encode()) & 0x7FFFFFFF)
    if kind == "ln_weight":
        return 1.0 + 0.02 * torch.randn(shape, generator=g)
    if kind == "exp_avg":
        return 1e-4 * torch.randn(shape, generator=g)
    if kind == "exp_avg_sq":
        return 1e-7 * torch.rand(shape, generator=g) + 1e-9
    return 0.02 * torch.randn(shape, generator=g)


def ckpt_model_state(named_shapes):
    sd = {}
    for k, shape in named_shapes:
        kind = "ln_weight" if k.endswith("layer_norm.weight") else "w"
        sd[k] = keyed_tensor("model/" + k, shape, kind)
    return sd


def ckpt_batches(n_steps, T, B, seed):
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n_steps):
        tok = rng.randint(2, 560, size=(T + 1, B))
        out.append((torch.from_numpy(tok[:-1]).long(), torch.from_numpy(tok[1:]).long(), torch.from_numpy(rng.rand(B) < 0.25)))
    return out


def gen_checkpoint_case(name):
    import io
    import json
    _stub_modules()
    from commu.model.dataset import BaseVocab            # the class the reference pickles into its checkpoints
    from commu.model.model import MemTransformerLM
    h = CKPT_HYPER
    torch.manual_seed(0)
    model = MemTransformerLM(ref_cfg(**CKPT_CFG), Vocab(729))
    names = [(k, tuple(p.shape)) for k, p in model.named_parameters()]
    sd0 = ckpt_model_state(names)
    with torch.no_grad():
        for k, p in model.named_parameters():
            p.copy_(sd0[k])
    model.train()

    def lam(step):
        if step == 0 and h["warmup"] == 0:
            return 1.0
        return max((h["warmup"] ** 0.5) / (step ** 0.5), h["lr_min"] / h["lr"]) if step > h["warmup"] else step / h["warmup"]
    opt = torch.optim.Adam(model.parameters(), lr=h["lr"], weight_decay=0.0)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lr_lambda=lam)
    # optimizer / scheduler state "after train_step steps": seeded moments, step counters, scheduler epoch
    osd = opt.state_dict()
    osd["state"] = {i: {"step": torch.tensor(float(h["train_step"])),
                        "exp_avg": keyed_tensor("opt/exp_avg/" + k, shp, "exp_avg"),
                        "exp_avg_sq": keyed_tensor("opt/exp_avg_sq/" + k, shp, "exp_avg_sq")}
                    for i, (k, shp) in enumerate(names)}
    opt.load_state_dict(osd)
    ssd = sched.state_dict()
    ssd["last_epoch"] = h["train_step"]
    ssd["_step_count"] = h["train_step"] + 1
    sched.load_state_dict(ssd)
    for g_ in opt.param_groups:
        g_["lr"] = h["lr"] * lam(h["train_step"])
    # ---- the reference's save_checkpoint dict (train.py:29-54), written and re-read through torch.save ----
    ckpt = {"model": model.state_dict(), "optimizer": opt.state_dict(), "train_step": h["train_step"],
            "scheduler": sched.state_dict(), "best_val_loss": 3.21, "vocab": BaseVocab(), "amp": None}
    buf = io.BytesIO()
    torch.save(ckpt, buf)
    buf.seek(0)
    back = torch.load(buf, weights_only=False)
    model.load_state_dict(back["model"], strict=False)            # model_initializer.py:46-47
    opt.load_state_dict(back["optimizer"])
    structure = {
        "model_keys": [[k, list(v.shape), str(v.dtype)] for k, v in back["model"].items()],
        "optimizer_param_groups": [{k: (v if not isinstance(v, tuple) else list(v)) for k, v in g_.items()}
                                   for g_ in back["optimizer"]["param_groups"]],
        "optimizer_state_keys": sorted(next(iter(back["optimizer"]["state"].values())).keys()),
        "optimizer_n_state": len(back["optimizer"]["state"]),
        "scheduler_keys": sorted(k for k in back["scheduler"].keys() if k != "lr_lambdas"),
        "top_level_keys": list(back.keys()),
        "vocab_class": type(back["vocab"]).__module__ + "." + type(back["vocab"]).__name__,
        "param_order": [k for k, _ in names],
        "file_bytes": buf.getbuffer().nbytes,
    }
    # ---- the reference continues training from it ----
    mems = [None] * h["chunks"]
    losses, lrs, gnorms = [], [], []
    for data, target, reset in ckpt_batches(h["n_steps"], CKPT_CFG["tgt_len"], h["B"], h["data_seed"]):
        model.zero_grad()
        tot = 0.0
        dc, tc, rc = torch.chunk(data, h["chunks"], 1), torch.chunk(target, h["chunks"], 1), torch.chunk(reset, h["chunks"], 0)
        for i in range(h["chunks"]):
            loss, mems[i] = model(dc[i].contiguous(), tc[i].contiguous(), rc[i].contiguous(), mems[i])
            loss = loss[tc[i] != 0].float().mean() / h["chunks"]
            tot += loss.item()
            loss.backward()
        gnorms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)))
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sched.step()
        losses.append(tot)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), structure=np.array(json.dumps(structure)),
                        losses=np.array(losses), lrs=np.array(lrs), gnorms=np.array(gnorms))
    print("wrote", name, "losses", losses, "lrs", lrs, "file MB", structure["file_bytes"] / 2 ** 20)


def main():
    sys.path.insert(0, REF)
    torch.set_num_threads(4)
    cfgA = dict(n_layer=2, n_head=4, d_model=64, d_inner=128, tgt_len=12, mem_len=16,
                same_length=False, clamp_len=-1)
    gen_forward_case("fwd_basic", cfgA, n_token=729, B=3, n_seg=3, seed=11, std=0.05,
                     reset_plan=[[0, 0, 0], [0, 1, 0], [1, 0, 0]])
    cfgB = dict(n_layer=2, n_head=6, d_model=60, d_inner=100, tgt_len=8, mem_len=24,
                same_length=True, clamp_len=20)
    gen_forward_case("fwd_samelen_dh10", cfgB, n_token=53, B=2, n_seg=5, seed=12, std=0.08,
                     reset_plan=[[0, 0], [0, 0], [1, 0], [0, 0], [0, 1]])
    cfgC = dict(n_layer=3, n_head=4, d_model=64, d_inner=128, tgt_len=1, mem_len=20,
                same_length=True, clamp_len=-1)
    gen_decode_case("decode_greedy", cfgC, n_token=97, B=2, n_ctx=6, n_new=40, seed=13, std=0.2)
    gen_sampler_case("sampler_probs", seed=14)
    gen_dataset_case("dataset_batches", seed=16)
    gen_teacher_case("teacher_forcing")
    gen_generate_case("generate_greedy")
    gen_checkpoint_case("checkpoint_default_shape")
    cfgE = dict(n_layer=2, n_head=2, d_model=32, d_inner=64, tgt_len=10, mem_len=10,
                same_length=False, clamp_len=-1)
    gen_train_case("train_steps", cfgE, n_token=61, B=4, chunks=2, n_steps=6, seed=15, std=0.05,
                   lr=0.004, warmup=3)


if __name__ == "__main__":
    main()
