"""Sampler surface of the reference's InferenceTask (commu/midi_generator/midi_inferrer.py:172-237)
on top of the native decode engine.  Method names, argument meaning and error behaviour follow the
reference so the host-side generation loop (`generate_sequence`, teacher forcing) can drive it
unchanged; the per-token arithmetic (1-token forward with cached memory, temperature / top-k /
top-p / wrong-token masking, multinomial draw) runs in libcommu_b200.so.

Kept quirks: token 0 is never sampled (Q4); `calc_probs` divides the cached
logits by the temperature IN PLACE, so a re-used logits tensor is divided again (Q3); a sampling
failure (all mass rejected) raises RuntimeError (Q5).
"""
import math
from typing import List, Tuple

import torch

from commu import _native as nv
from commu.engine.decode import DecodeEngine, DecodeState
from commu.preprocessor.encoder.event_tokens import TOKEN_OFFSET

try:
    from logger import logger
except ImportError:  # pragma: no cover
    import logging
    logger = logging.getLogger("ComMU")

_T = {k: v.value for k, v in TOKEN_OFFSET.__members__.items()}
POSITION_RESOLUTION = 128   # commu/preprocessor/utils/constants.py:25


class TeacherForceTask:
    """Chord teacher forcing of the reference's generation loop (midi_inferrer.py:16-169), restated as a
    queue of pending chords: every entry is (chord token, position token, is_inter) where is_inter marks a
    chord that does not sit on the first position of a bar.  Public method names are the reference's,
    because `generate_sequence` drives the task through them."""

    def __init__(self, input_data):
        self.input_data = input_data
        self.next_tokens_forced: List[int] = []
        self.wrong_tokens: List[int] = []
        self.no_sequence_appended = False
        self.is_incomplete = input_data.num_measures % 4 != 0
        self.incomplete_filled = not self.is_incomplete
        tokens, positions = input_data.chord_token_components.values()
        assert len(tokens) == len(positions), "Wrong Chord Length"
        self.pending = [(int(t), int(pos), int(pos) != _T["POSITION"]) for t, pos in zip(tokens, positions)]
        self.chord_length = len(self.pending)

    # views with the reference's attribute names
    chord_token = property(lambda self: [c[0] for c in self.pending])
    chord_position = property(lambda self: [c[1] for c in self.pending])
    inter_chord_flags = property(lambda self: [c[2] for c in self.pending])

    # ---- predicates -------------------------------------------------------------------------------
    def check_remnant_chord(self):
        return len(self.pending) > 0

    def check_first_position(self, seq):
        return self.incomplete_filled and seq[-1] == _T["BAR"]

    def check_length_fit(self):
        return self.chord_length == int(self.input_data.num_measures // 4 * 4)

    def check_position_fit(self, seq):
        return seq[-2] == _T["BAR"] and seq[-1] == _T["POSITION"]

    def _chord_due(self, seq):
        return self.check_remnant_chord() and self.incomplete_filled

    def check_one_chord_per_bar_case(self, seq):
        return self._chord_due(seq) and self.check_length_fit() and self.check_position_fit(seq)

    def check_mul_chord_per_bar_case(self, seq):
        if not self._chord_due(seq) or self.check_length_fit():
            return False
        if self.check_position_fit(seq):
            return True
        _, pos, inter = self.pending[0]
        return seq[-1] == pos and inter

    def check_chord_position_passed(self, token):
        if not self.check_remnant_chord():
            return False
        _, pos, inter = self.pending[0]
        passed = pos < token < _T["POSITION"] + POSITION_RESOLUTION or token == _T["BAR"]
        return inter and passed

    def check_wrong_chord_token_generated(self, token):
        return _T["CHORD_START"] <= token <= _T["CHORD_END"]

    def check_wrong_eos_generated(self, token):
        return self.check_remnant_chord() and token == _T["EOS"]

    def check_wrong_bar_token_generated(self, token):
        return not self.check_remnant_chord() and token == _T["BAR"]

    # ---- actions ----------------------------------------------------------------------------------
    def teach_first_position(self):
        self.next_tokens_forced.append(_T["POSITION"])

    def teach_chord_token(self):
        tok, _, _ = self.pending.pop(0)
        self.next_tokens_forced.append(tok)
        self.wrong_tokens = []

    def teach_chord_position(self):
        self.next_tokens_forced.append(self.pending[0][1])
        self.wrong_tokens = []

    def teach_wrong_chord_token(self, wrong_token):
        self.no_sequence_appended = True
        self.wrong_tokens.append(wrong_token)

    def teach_remnant_chord(self):
        _, pos, inter = self.pending[0]
        self.next_tokens_forced.append(pos if inter else _T["BAR"])

    def teach_eos(self):
        self.next_tokens_forced.append(_T["EOS"])

    def validate_teacher_forced_sequence(self, seq):
        n_bars = seq.count(_T["BAR"])
        n_chords = sum(1 for t in seq if _T["CHORD_START"] <= t <= _T["CHORD_END"])
        if self.pending:
            raise Exception(f"remnant chord length: {len(self.pending)} \nerror in teacher forcing")
        if n_bars != int(math.ceil(self.input_data.num_measures)):
            raise Exception(f"bar length: {n_bars} \nerror in bar length")
        if n_chords != self.chord_length:
            raise Exception(f"num_chord: {n_chords} vs {self.chord_length} \nerror in chord length")
        logger.info(f"correct_length: {n_bars}")
        logger.info(seq)


class InferenceTask:
    def __init__(self, device: torch.device, precision: str = "fp32", seed: int = 0):
        self.device = device
        self.precision = precision
        self.seed = seed
        self._draws = 0
        self.engine = None

    def __call__(self, model, input_data, inference_cfg):
        self.model = model
        self.input_data = input_data
        self.inference_cfg = inference_cfg
        self.engine = DecodeEngine(model, batch=1, mem_len=model.mem_len, same_length=model.same_length,
                                   precision=self.precision)

    def _tok(self, t):
        return torch.tensor([t], dtype=torch.int64, device=self.device)

    def init_seq_and_mems(self, encoded_meta: List[int], num_conditional_tokens: int):
        seq = [0]
        state = DecodeState()
        for t in seq + encoded_meta[: num_conditional_tokens - 1]:
            _, state = self.engine.step(self._tok(t), state)
        return seq + encoded_meta[:num_conditional_tokens], state

    def calc_logits_and_mems(self, seq: List[int], mems) -> Tuple[torch.Tensor, DecodeState]:
        logits, state = self.engine.step(self._tok(seq[-1]), mems)
        return logits[0, 1:].clone(), state          # drop token 0 (reference :206)

    def calc_probs(self, logits):
        V = logits.shape[0] + 1
        full = torch.empty(1, V, device=logits.device)
        full[0, 0] = 0.0
        temp = float(self.input_data.temperature)
        if temp != 0:
            logits /= temp                             # in place, like the reference (:216)
            full[0, 1:] = logits
            temp = 1.0                                 # already applied
        else:
            full[0, 1:] = logits
        probs = torch.empty(1, V, device=logits.device)
        nv.call("commu_sample", full, V, 1, V, temp, 0, 0.0, None, 0, 0, None, probs, V, None)
        return probs[0]

    def apply_sampling(self, probs, wrong_tokens):
        V = probs.shape[0]
        wrong = None
        if wrong_tokens:
            wrong = torch.zeros(1, V, dtype=torch.uint8, device=probs.device)
            wrong[0, list(wrong_tokens)] = 1
        # feed log-probabilities back through the sampler with temperature 1: softmax(log p) == p
        lg = torch.log(probs).unsqueeze(0).contiguous()
        out = torch.empty(1, V, device=probs.device)
        top_p = float(getattr(self.input_data, "top_p", 0.0) or 0.0)
        nv.call("commu_sample", lg, V, 1, V, 1.0, int(self.input_data.top_k), top_p, wrong, 0, 0, None, out, V, None)
        probs.copy_(out[0])
        return probs

    def infer_token(self, probs):
        if not bool(torch.isfinite(probs).all()) or float(probs.sum()) <= 0:
            raise RuntimeError("probability tensor contains either `inf`, `nan` or element < 0")
        lg = torch.log(probs).unsqueeze(0).contiguous()
        V = probs.shape[0]
        tok = torch.empty(1, dtype=torch.int64, device=probs.device)
        self._draws += 1
        nv.call("commu_sample", lg, V, 1, V, 1.0, 0, 0.0, None, self.seed, self._draws, tok, None, V, None)
        return int(tok.item())

    # ------------------------------------------------------------------------------------------------
    # host-side generation loop (reference midi_inferrer.py:239-354); quirks Q1-Q5 of SURVEY.md 3.2 kept
    # ------------------------------------------------------------------------------------------------
    def generate_sequence(self, seq, mems):
        logits = None
        teacher = TeacherForceTask(self.input_data)
        first = True
        for _ in range(self.inference_cfg.GENERATION.generation_length):
            if seq[-1] == _T["EOS"]:
                break
            if teacher.next_tokens_forced:                       # forced tokens are fed right away (Q2)
                seq.append(teacher.next_tokens_forced.pop(0))
                logits, mems = self.calc_logits_and_mems(seq, mems)
                continue
            if teacher.no_sequence_appended:                     # rejected token: re-use the logits (Q3)
                assert logits is not None
                teacher.no_sequence_appended = False
            elif first:                                          # the returned memory is dropped once (Q1)
                logits, _ = self.calc_logits_and_mems(seq, mems)
                first = False
            else:
                logits, mems = self.calc_logits_and_mems(seq, mems)
            probs = self.apply_sampling(self.calc_probs(logits), teacher.wrong_tokens)
            if not teacher.incomplete_filled:
                teacher.incomplete_filled = seq.count(_T["BAR"]) > 1
            if teacher.check_first_position(seq):
                teacher.teach_first_position()
                continue
            if teacher.check_one_chord_per_bar_case(seq) or teacher.check_mul_chord_per_bar_case(seq):
                teacher.teach_chord_token()
                continue
            try:
                token = self.infer_token(probs)
            except RuntimeError as e:
                logger.error(f"Sampling Error: {e}")
                seq = None
                break
            if teacher.check_chord_position_passed(token):
                teacher.teach_chord_position()
            elif teacher.check_wrong_chord_token_generated(token):
                teacher.teach_wrong_chord_token(token)
            elif teacher.check_wrong_eos_generated(token):
                teacher.teach_remnant_chord()
            elif teacher.check_wrong_bar_token_generated(token):
                teacher.teach_eos()
            else:
                seq.append(token)
        try:
            teacher.validate_teacher_forced_sequence(seq)
        except Exception as err:  # noqa: BLE001 - same catch-all as the reference
            logger.error(err)
            seq = None
        return seq

    def validate_generated_sequence(self, seq: List[int]) -> bool:
        notes = 0
        for k, tok in enumerate(seq):
            if k + 2 > len(seq) - 1:
                break
            if _T["NOTE_VELOCITY"] <= tok < _T["CHORD_START"]:
                if (_T["POSITION"] <= seq[k - 1] < _T["BPM"] and _T["PITCH"] <= seq[k + 1] < _T["NOTE_VELOCITY"]
                        and _T["NOTE_DURATION"] <= seq[k + 2] < _T["POSITION"]):
                    notes += 1
        return notes > 0

    # ------------------------------------------------------------------------------------------------
    # batched generation: the SAME loop, one coroutine per sequence, lock-stepped on the
    # batch decode engine.  Every sequence asks for a model step exactly where the reference calls
    # calc_logits_and_mems and for a token where it calls infer_token; all sequences of a batch receive their
    # step from ONE launch sequence (one token each), their samples from one sampler launch with per-sequence
    # wrong-token masks on the device, and the host sees one [B] token copy per wave instead of one .item() per
    # token and sequence.  Quirks kept: Q1 (the first step's memory is dropped: the shared ring state is rewound
    # once), Q2 (a forced token is fed again by the next ordinary iteration), Q3 (re-used logits are divided by the
    # temperature again: the sampler gets temperature ** k), Q4 (token 0 never sampled), Q5 (sampling failure
    # discards the sequence).
    # ------------------------------------------------------------------------------------------------
    def _sequence_steps(self, seq):
        """generate_sequence (reference midi_inferrer.py:239-320) as a coroutine.  Yields ("step", token, keep) /
        ("sample", k, wrong_tokens); returns the finished sequence or None."""
        teacher = TeacherForceTask(self.input_data)
        first, have_logits, ndiv = True, False, 0
        for _ in range(self.inference_cfg.GENERATION.generation_length):
            if seq[-1] == _T["EOS"]:
                break
            if teacher.next_tokens_forced:
                seq.append(teacher.next_tokens_forced.pop(0))
                yield ("step", seq[-1], True)
                have_logits, ndiv = True, 0
                continue
            if teacher.no_sequence_appended:
                assert have_logits
                teacher.no_sequence_appended = False
            elif first:
                yield ("step", seq[-1], False)
                first, have_logits, ndiv = False, True, 0
            else:
                yield ("step", seq[-1], True)
                have_logits, ndiv = True, 0
            ndiv += 1                                            # calc_probs: logits /= temperature, in place
            if not teacher.incomplete_filled:
                teacher.incomplete_filled = seq.count(_T["BAR"]) > 1
            if teacher.check_first_position(seq):
                teacher.teach_first_position()
                continue
            if teacher.check_one_chord_per_bar_case(seq) or teacher.check_mul_chord_per_bar_case(seq):
                teacher.teach_chord_token()
                continue
            token = yield ("sample", ndiv, list(teacher.wrong_tokens))
            if token < 0:                                        # all mass rejected (reference: RuntimeError)
                logger.error("Sampling Error: probability tensor contains either `inf`, `nan` or element < 0")
                return None
            if teacher.check_chord_position_passed(token):
                teacher.teach_chord_position()
            elif teacher.check_wrong_chord_token_generated(token):
                teacher.teach_wrong_chord_token(token)
            elif teacher.check_wrong_eos_generated(token):
                teacher.teach_remnant_chord()
            elif teacher.check_wrong_bar_token_generated(token):
                teacher.teach_eos()
            else:
                seq.append(token)
        try:
            teacher.validate_teacher_forced_sequence(seq)
        except Exception as err:  # noqa: BLE001 - same catch-all as the reference
            logger.error(err)
            return None
        return seq

    @torch.no_grad()
    def generate_batch(self, encoded_meta: List[int], n: int):
        """Runs n sequences of the same metadata together.  Returns a list of n entries (sequence or None)."""
        eng = self._batch_engine(n)
        B, V, dev = eng.B, eng.V, self.device
        temp = float(self.input_data.temperature)
        top_k = int(self.input_data.top_k)
        top_p = float(getattr(self.input_data, "top_p", 0.0) or 0.0)
        ctx = [0] + list(encoded_meta)
        state = DecodeState()
        for t in ctx[:-1]:                                       # init_seq_and_mems: memory of [0] + meta[:-1]
            _, state = eng.step(torch.full((B,), t, dtype=torch.int64, device=dev), state)
        gens = [self._sequence_steps(list(ctx)) for _ in range(n)]
        done = [None] * n
        live = {}

        def advance(row, value=None):
            """Runs sequence `row` up to its next request (or to its end)."""
            try:
                live[row] = gens[row].send(value) if row in live else next(gens[row])
            except StopIteration as fin:
                live.pop(row, None)
                done[row] = fin.value
        for row in range(n):
            advance(row)
        tok_h = torch.zeros(B, dtype=torch.int64).pin_memory()
        wrong = torch.zeros(B, V, dtype=torch.uint8, device=dev)
        toks_d = torch.empty(B, dtype=torch.int64, device=dev)
        while live:
            # ---- one model step for every live sequence ----
            keeps = {req[2] for req in live.values()}
            assert all(req[0] == "step" for req in live.values()) and len(keeps) == 1
            tok_h.zero_()
            for row, req in live.items():
                tok_h[row] = req[1]
            logits, new_state = eng.step(tok_h.to(dev), state)
            if keeps.pop():
                state = new_state                                # (Q1: the first step's memory is dropped)
            for row in list(live):
                advance(row)
            # ---- sampling waves: first draws together, re-draws after a rejected token as they come ----
            while any(req[0] == "sample" for req in live.values()):
                rows = [r for r, req in live.items() if req[0] == "sample"]
                k = min(live[r][1] for r in rows)
                rows = [r for r in rows if live[r][1] == k]
                wrong.zero_()
                for r in rows:
                    if live[r][2]:
                        wrong[r, live[r][2]] = 1
                self._draws += 1
                nv.call("commu_sample", logits, logits.stride(0), B, V, temp ** k if temp != 0 else 0.0, top_k, top_p,
                        wrong, self.seed, self._draws, toks_d, None, V, None)
                got = toks_d.cpu()
                for r in rows:
                    advance(r, int(got[r]))
        return done

    def _batch_engine(self, n):
        if n > 64:
            raise ValueError("generate_batch: at most 64 sequences per batch")
        eng = getattr(self, "_beng", None)
        if eng is None or eng.B < n:
            eng = DecodeEngine(self.model, batch=n, mem_len=self.model.mem_len, same_length=self.model.same_length,
                               precision=self.precision)
            self._beng = eng
        return eng

    def execute_batched(self, encoded_meta, max_rounds=None) -> List[List[int]]:
        """execute() with the sequences of a round generated together; failed sequences (teacher-forcing
        validation, empty sequence, sampling failure) are regenerated in the next round, like the reference."""
        want = int(self.input_data.num_generate)
        sequences, rounds = [], 0
        while len(sequences) < want and (max_rounds is None or rounds < max_rounds):
            rounds += 1
            n = min(want - len(sequences), 64)
            logger.info("Generating %d sequence(s) together (round %d)" % (n, rounds))
            for seq in self.generate_batch(encoded_meta, n):
                if seq is None:
                    continue
                if not self.validate_generated_sequence(seq):
                    logger.error("Empty sequence generated")
                    continue
                sequences.append(seq)
        return sequences[:want]

    def execute(self, encoded_meta) -> List[List[int]]:
        n_cond = len(encoded_meta)
        sequences = []
        while len(sequences) != self.input_data.num_generate:
            with torch.no_grad():
                logger.info("Generating the idx: " + str(len(sequences) + 1))
                seq, mems = self.init_seq_and_mems(encoded_meta, n_cond)
                seq = self.generate_sequence(seq, mems)
                if seq is None:
                    continue
                if not self.validate_generated_sequence(seq):
                    logger.error("Empty sequence generated")
                    continue
            sequences.append(seq)
        return sequences
