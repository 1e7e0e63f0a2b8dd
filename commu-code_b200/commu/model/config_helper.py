"""Configuration surface of the reference (commu/model/config_helper.py) without the yacs
dependency: same section / field names and defaults, `freeze()` / `defrost()` / `str(cfg)`, plus an
override hook the reference lacks (`COMMU_CFG_OPTS="MODEL.num_layers=12,TRAIN.tgt_length=2048"` or
`apply_overrides(cfg, {...})`) because the benchmark shapes cannot be expressed otherwise."""
import ast
import os


class CfgNode(dict):
    """Minimal attribute-dict with the slice of the yacs API the reference touches."""

    _FROZEN = "__frozen__"

    def __init__(self, init=None):
        super().__init__()
        object.__setattr__(self, CfgNode._FROZEN, False)
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if object.__getattribute__(self, CfgNode._FROZEN):
            raise AttributeError("Attempted to set %s on a frozen CfgNode" % name)
        self[name] = value

    def _set_frozen(self, flag):
        object.__setattr__(self, CfgNode._FROZEN, flag)
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def freeze(self):
        self._set_frozen(True)

    def defrost(self):
        self._set_frozen(False)

    def is_frozen(self):
        return object.__getattribute__(self, CfgNode._FROZEN)

    def clone(self):
        return CfgNode({k: (v.clone() if isinstance(v, CfgNode) else v) for k, v in self.items()})

    def __str__(self):
        def dump(node, indent):
            out = []
            for k in sorted(node.keys()):
                v = node[k]
                if isinstance(v, CfgNode):
                    out.append(" " * indent + "%s:" % k)
                    out.extend(dump(v, indent + 2))
                else:
                    out.append(" " * indent + "%s: %s" % (k, v))
            return out
        return "\n".join(dump(self, 0))


_TRAINING_DEFAULTS = {
    "INITIALIZER": {"base_init": 0.01, "embed_init": 0.01},
    "EVALUATE": {"batch_size": 10, "tgt_length": 128, "mem_length": 2048},
    "MODEL": {"num_layers": 6, "num_heads": 10, "units": 500, "inner_size": 1000, "dropout": 0.1,
              "attention_dropout": 0.1, "clamp_len": -1, "same_length": False},
    "TRAIN": {"batch_size": 256, "batch_chunk": 4, "tgt_length": 128, "mem_length": 1024, "seed": 1111,
              "lr": 0.004, "lr_min": 0.0001, "warmup_step": 100, "clip": 1.0, "max_step": 20000,
              "log_interval": 100, "eval_interval": 1000, "weight_decay": 0.0},
}
_INFERENCE_DEFAULTS = {
    "MODEL": {"memory_length": 4146, "device": "gpu"},
    "SAMPLING": {"threshold": 32.0, "temperature": 0.95},
    "GENERATION": {"generation_length": 4096},
}


def apply_overrides(cfg, overrides):
    """overrides: {"SECTION.field": value}.  Works on an unfrozen cfg."""
    for key, val in overrides.items():
        sec, field = key.split(".")
        if sec not in cfg or field not in cfg[sec]:
            raise KeyError("unknown config field %s" % key)
        cfg[sec][field] = val
    return cfg


def _env_overrides():
    spec = os.environ.get("COMMU_CFG_OPTS", "").strip()
    out = {}
    for item in filter(None, (s.strip() for s in spec.split(","))):
        k, v = item.split("=", 1)
        try:
            out[k.strip()] = ast.literal_eval(v.strip())
        except (ValueError, SyntaxError):
            out[k.strip()] = v.strip()
    return out


def get_default_cfg_training(overrides=None):
    cfg = CfgNode(_TRAINING_DEFAULTS)
    apply_overrides(cfg, _env_overrides())
    if overrides:
        apply_overrides(cfg, overrides)
    cfg.freeze()
    return cfg


def get_default_cfg_inference():
    cfg = CfgNode(_INFERENCE_DEFAULTS)
    cfg.freeze()
    return cfg
