"""Generation entry point with the reference's command line (generate.py of POZAlabs/ComMU-code:
`--checkpoint_dir`, `--output_dir`, the 11 metadata flags, `--chord_progression`, `--num_generate --top_k
--temperature`) and the reference's generation semantics: the sequences come from
`InferenceTask.execute` - chord teacher forcing, wrong-token rejection, the validate / regenerate loop
(commu/midi_generator/midi_inferrer.py:239-354 of the reference) - running on the native KV-cache decode
engine; `num_generate` sequences are decoded TOGETHER (one batched token step and one sampler launch per
step, `InferenceTask.execute_batched`; `--sequential` selects the reference's one-by-one loop).

Out of scope: the metadata -> token encoder, the chord-name ->
chord-token table and the tokens -> MIDI writer of the reference are host-side utilities that are not mirrored
here.  The ALREADY-ENCODED values are therefore passed instead:
    --meta_tokens       the 11 encoded meta tokens (what `MetaEncoder` would produce from the musical flags)
    --chord_tokens      the chord tokens of `TransXlInputData.chord_token_components["chord_token"]`
    --chord_positions   ... and ["chord_position"] (432 = first position of a bar)
    --num_measures      as in the reference
A textual `--chord_progression` (or musical flags without `--meta_tokens`) is rejected loudly instead of being
ignored.  Output: `<output_dir>/generated_<i>.npy` with the full token sequence ([0] + meta + events).
Additions: `--top_p`, `--precision fp32|bf16`, `--seed`, `--max_rounds`.
"""
import argparse
import os
import sys
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "commu-code_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from commu.midi_generator.midi_inferrer import InferenceTask  # noqa: E402
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
    ap.add_argument("--chord_tokens", type=str, default=None, help="encoded chord tokens, comma separated")
    ap.add_argument("--chord_positions", type=str, default=None, help="encoded chord positions, comma separated")
    ap.add_argument("--generation_length", type=int, default=None, help="overrides GENERATION.generation_length (4096)")
    ap.add_argument("--max_rounds", type=int, default=None, help="stop regenerating failed sequences after this many rounds")
    ap.add_argument("--sequential", action="store_true", help="the reference's one-sequence-at-a-time loop")
    ap.add_argument("--precision", type=str, default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--opts", type=str, default="", help="training-config overrides used to build the model")
    return ap.parse_args(argv)


def _ints(s):
    return [int(t) for t in s.split(",") if t.strip() != ""]


def build_input_data(args):
    """The fields of the reference's TransXlInputData (commu/midi_generator/container.py:17-62) the generation loop
    reads, from already-encoded values."""
    if not args.meta_tokens:
        given = {f: getattr(args, f) for f in META_FLAGS if getattr(args, f) is not None}
        raise SystemExit("generate.py: metadata encoding is outside this repository's scope; pass the 11 encoded "
                         "meta tokens with --meta_tokens (got musical flags: %s)" % given)
    if args.chord_progression and not args.chord_tokens:
        raise SystemExit("generate.py: --chord_progression takes chord NAMES, whose encoder is outside this "
                         "repository's scope; pass the encoded chords with --chord_tokens / --chord_positions "
                         "(TransXlInputData.chord_token_components of the reference) - refusing to ignore it")
    if args.num_measures is None:
        raise SystemExit("generate.py: --num_measures is required (teacher forcing validates the bar count)")
    ctok = _ints(args.chord_tokens or "")
    cpos = _ints(args.chord_positions or "")
    if len(ctok) != len(cpos):
        raise SystemExit("generate.py: --chord_tokens and --chord_positions must have the same length")
    meta = _ints(args.meta_tokens)
    data = SimpleNamespace(num_measures=float(args.num_measures), temperature=args.temperature, top_k=args.top_k,
                           top_p=args.top_p, num_generate=args.num_generate,
                           chord_token_components={"chord_token": ctok, "chord_position": cpos})
    return meta, data


def main(argv=None):
    args = parse_args(argv)
    meta, input_data = build_input_data(args)
    device = torch.device("cuda", 0)
    overrides = {}
    for item in filter(None, (s.strip() for s in args.opts.split(","))):
        k, v = item.split("=", 1)
        overrides[k] = int(v) if v.lstrip("-").isdigit() else (v == "True" if v in ("True", "False") else float(v))
    init = ModelInitializeTask(args, map_location=device, device=device)
    model = init.execute(overrides)
    inference_cfg = init.inference_cfg
    if args.generation_length:
        inference_cfg = SimpleNamespace(GENERATION=SimpleNamespace(generation_length=args.generation_length))
    task = InferenceTask(device, precision=args.precision, seed=args.seed)
    task(model=model, input_data=input_data, inference_cfg=inference_cfg)
    logger.info("Generating %d sequence(s), %d chords to teach" % (args.num_generate, len(input_data.chord_token_components["chord_token"])))
    if args.sequential:
        sequences = task.execute(meta)
    else:
        sequences = task.execute_batched(meta, max_rounds=args.max_rounds)
    os.makedirs(args.output_dir, exist_ok=True)
    for b, seq in enumerate(sequences):
        np.save(os.path.join(args.output_dir, "generated_%d.npy" % b), np.asarray(seq, dtype=np.int64))
        logger.info("sequence %d: %d tokens" % (b, len(seq)))
    if len(sequences) < args.num_generate:
        logger.error("only %d of %d sequences passed validation within %s rounds" % (len(sequences), args.num_generate, args.max_rounds))
    return sequences


if __name__ == "__main__":
    main()
