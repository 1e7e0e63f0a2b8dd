"""Sampler surface of the reference's InferenceTask (commu/midi_generator/midi_inferrer.py:172-237)
on top of the native decode engine.  Method names, argument meaning and error behaviour follow the
reference so the host-side generation loop (`generate_sequence`, teacher forcing) can drive it
unchanged; the per-token arithmetic (1-token forward with cached memory, temperature / top-k /
top-p / wrong-token masking, multinomial draw) runs in libcommu_b200.so.

Kept quirks: token 0 is never sampled (Q4); `calc_probs` divides the cached
logits by the temperature IN PLACE, so a re-used logits tensor is divided again (Q3); a sampling
failure (all mass rejected) raises RuntimeError (Q5).
"""
from typing import List, Tuple

import torch

from commu import _native as nv
from commu.engine.decode import DecodeEngine, DecodeState


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
