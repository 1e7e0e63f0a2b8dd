"""Checkpoint -> ready-to-decode model, with the reference's behaviour
(commu/midi_generator/model_initializer.py:13-56): the model is ALWAYS built from the default training
config with same_length=True (the run's config.yml is never read, quirk Q7), the checkpoint's
"model" entry is loaded with strict=False, eval mode, reset_length(1, memory_length)."""
from pathlib import Path

import torch

from commu.model.config_helper import get_default_cfg_inference, get_default_cfg_training
from commu.model.dataset import BaseVocab
from commu.model.model import MemTransformerLM


class ModelInitializeTask:
    def __init__(self, model_args, map_location, device):
        self.model_args = model_args
        self.map_location = map_location
        self.device = device
        self.inference_cfg = get_default_cfg_inference()

    def load_checkpoint_fp(self):
        ckpt = getattr(self.model_args, "checkpoint_dir", None)
        if not ckpt:
            raise FileNotFoundError("--checkpoint_dir is required (the reference's fallback path "
                                    "inference_cfg.MODEL.model_directory does not exist either)")
        fp = Path(ckpt)
        return fp, fp.parent / "config.yml"

    def initialize_training_cfg(self, overrides=None):
        ov = {"MODEL.same_length": True}
        ov.update(overrides or {})
        return get_default_cfg_training(ov)

    def initialize_model(self, training_cfg, model_fp):
        model = MemTransformerLM(training_cfg, BaseVocab())
        ckpt = torch.load(model_fp, map_location=self.map_location, weights_only=False)  # pickled BaseVocab inside
        model.load_state_dict(ckpt["model"], strict=False)
        model = model.to(self.device)
        model.eval()
        model.reset_length(1, self.inference_cfg.MODEL.memory_length)
        return model

    def execute(self, overrides=None):
        model_fp, _ = self.load_checkpoint_fp()
        return self.initialize_model(self.initialize_training_cfg(overrides), model_fp)
