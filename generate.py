"""Generation entry point with the reference's command line (generate.py of POZAlabs/ComMU-code:
`--checkpoint_dir`, `--output_dir`, the 11 metadata flags, `--num_generate --top_k --temperature`)
on the native KV-cache decode engine.  Additions: `--top_p`, `--precision fp32|bf16`,
`--meta_tokens a,b,...` (the 11 already-encoded meta tokens) and `--max_new_tokens`.

The metadata -> token encoder and the tokens -> MIDI writer of the reference are host-side utilities
outside this repository's scope; without `--meta_tokens` the musical flags are
accepted and reported but cannot be encoded here, and the output is the generated token ids
(`<output_dir>/generated_<i>.npy`).  `num_generate` sequences are decoded TOGETHER as one batch.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from commu.engine.decode import DecodeEngine  # noqa: E402
from commu.midi_generator.model_initializer import ModelInitializeTask  # noqa: E402
from logger import logger  # noqa: E402

META_FLAGS = ["bpm", "audio_key", "time_signature", "pitch_range", "num_measures", "inst", "genre",
              "min_velocity", "max_velocity", "track_role", "rhythm"]


def parse_args(argv=None):
    ap = argparse.ArgumentParser(description="ComMU generation (B200-native)")
    ap.add_argument("--checkpoint_dir", type=str, required=True)
    ap.add_argument("--output_dir", type=str, required=True)
    for f in META_FLAGS:
        ap.add_argument("--" + f, type=str, default=None)
    ap.add_argument("--chord_progression", type=str, default=None)
    ap.add_argument("--num_generate", type=int, default=1)
    ap.add_argument("--top_k", type=int, default=32)
    ap.add_argument("--temperature", type=float, default=0.95)
    ap.add_argument("--top_p", type=float, default=0.0)
    ap.add_argument("--meta_tokens", type=str, default=None, help="11 encoded meta tokens, comma separated")
    ap.add_argument("--max_new_tokens", type=int, default=4096)
    ap.add_argument("--precision", type=str, default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--opts", type=str, default="", help="training-config overrides used to build the model")
    return ap.parse_args(argv)


def main(argv=None):
    args = parse_args(argv)
    if not args.meta_tokens:
        given = {f: getattr(args, f) for f in META_FLAGS if getattr(args, f) is not None}
        raise SystemExit("generate.py: metadata encoding is outside this repository's scope; pass the 11 encoded "
                         "meta tokens with --meta_tokens (got musical flags: %s)" % given)
    meta = [int(t) for t in args.meta_tokens.split(",")]
    device = torch.device("cuda", 0)
    overrides = {}
    for item in filter(None, (s.strip() for s in args.opts.split(","))):
        k, v = item.split("=", 1)
        overrides[k] = int(v) if v.lstrip("-").isdigit() else (v == "True" if v in ("True", "False") else float(v))
    model = ModelInitializeTask(args, map_location=device, device=device).execute(overrides)
    eng = DecodeEngine(model, batch=args.num_generate, mem_len=model.mem_len, same_length=model.same_length,
                       precision=args.precision)
    # reference context: [0] + meta tokens (midi_inferrer.py:186-197)
    ctx = torch.tensor([[0] + meta] * args.num_generate, dtype=torch.int64, device=device).t().contiguous()
    logger.info("Generating %d sequence(s) of up to %d tokens" % (args.num_generate, args.max_new_tokens))
    toks = eng.generate(ctx, args.max_new_tokens, temperature=args.temperature, top_k=args.top_k,
                        top_p=args.top_p, seed=args.seed).cpu().numpy()
    os.makedirs(args.output_dir, exist_ok=True)
    for b in range(args.num_generate):
        seq = toks[:, b]
        eos = np.nonzero(seq == 1)[0]
        if len(eos):
            seq = seq[: eos[0] + 1]
        full = np.concatenate(([0], meta, seq))
        np.save(os.path.join(args.output_dir, "generated_%d.npy" % b), full)
        logger.info("sequence %d: %d tokens" % (b, len(full)))


if __name__ == "__main__":
    main()
