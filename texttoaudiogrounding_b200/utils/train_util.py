"""Config / registry surface of the reference's strong-supervision runner
(reference utils/train_util.py:120-137 get_obj_from_str / init_obj_from_str, :153-194 config
loading with `inherit_from`, python_scripts/training/run_strong.py:71-89 get_model)."""
from __future__ import annotations

import importlib
import os
import sys
from typing import Dict

import yaml


def get_obj_from_str(string: str, reload: bool = False):
    module, cls = string.rsplit(".", 1)
    if reload:
        importlib.reload(importlib.import_module(module))
    return getattr(importlib.import_module(module, package=None), cls)


def init_obj_from_str(config: Dict, **kwargs):
    obj_args = config["args"].copy()
    obj_args.update(kwargs)
    for k in config:
        if k not in ["type", "args"] and isinstance(config[k], dict) and k not in kwargs:
            obj_args[k] = init_obj_from_str(config[k])
    cls = get_obj_from_str(config["type"])
    return cls(**obj_args)


def merge_a_into_b(a: Dict, b: Dict) -> None:
    for k, v in a.items():
        if isinstance(v, dict) and k in b and isinstance(b[k], dict):
            merge_a_into_b(v, b[k])
        else:
            b[k] = v


def load_config(config_file: str) -> Dict:
    with open(config_file, "r") as reader:
        config = yaml.load(reader, Loader=yaml.FullLoader)
    if "inherit_from" in config:
        base_config_file = config["inherit_from"]
        base_config_file = os.path.join(os.path.dirname(config_file), base_config_file)
        assert not os.path.samefile(config_file, base_config_file), "inherit from itself"
        base_config = load_config(base_config_file)
        del config["inherit_from"]
        merge_a_into_b(config, base_config)
        return base_config
    return config


def parse_config_or_kwargs(config_file: str, **kwargs) -> Dict:
    """YAML + dotted ``a.b.c=value`` keyword overrides (the reference renders CLI kwargs to TOML
    and merges them; the observable result — nested-key override — is the same)."""
    config = load_config(config_file)
    for key, value in kwargs.items():
        node = config
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = value
    return config


def get_model(config: Dict, print_fn=sys.stdout.write):
    """run_strong.py:71-89: build every sub-module of cfg['model'] (optionally loading its
    `pretrained`), then the top-level type with them as kwargs."""
    from ..models.base import merge_matched_keys  # noqa: F401  (kept for parity of surface)
    import torch
    kwargs = {}
    for k in config["model"]:
        if k not in ["type", "args", "pretrained"]:
            sub_model = init_obj_from_str(config["model"][k])
            if "pretrained" in config["model"][k]:
                path = config["model"][k]["pretrained"]
                if os.path.exists(path):
                    if hasattr(sub_model, "load_pretrained"):
                        sub_model.load_pretrained(path, print_fn)
                    else:
                        sd = torch.load(path, map_location="cpu")
                        sd = sd.get("model", sd)
                        own = sub_model.state_dict()
                        own.update({k2: v for k2, v in sd.items() if k2 in own and own[k2].shape == v.shape})
                        sub_model.load_state_dict(own)
                else:
                    print_fn(f"pretrained {path} not exist!")
            kwargs[k] = sub_model
    return init_obj_from_str(config["model"], **kwargs)


_ALIASES = {
    "models": "texttoaudiogrounding_b200.models",
    "models.audio_encoder": "texttoaudiogrounding_b200.models.audio_encoder",
    "models.text_encoder": "texttoaudiogrounding_b200.models.text_encoder",
    "models.match": "texttoaudiogrounding_b200.models.match",
    "models.audio_text_model": "texttoaudiogrounding_b200.models.audio_text_model",
    "models.utils": "texttoaudiogrounding_b200.models.utils",
    "models.base": "texttoaudiogrounding_b200.models.base",
    "models.align": "texttoaudiogrounding_b200.models.align",
    "models.sim_pooling": "texttoaudiogrounding_b200.models.sim_pooling",
    "models.cross_encoder": "texttoaudiogrounding_b200.models.cross_encoder",
    "models.hf_modeling_grounding": "texttoaudiogrounding_b200.models.hf_modeling_grounding",
    "losses": "texttoaudiogrounding_b200.losses",
}


def install_as_reference_modules() -> None:
    """Register ``models.*`` / ``losses`` aliases in sys.modules so that an unmodified
    run_strong.py-dialect YAML (``type: models.audio_encoder.Cnn8Rnn`` ...) instantiates the
    B200 classes."""
    for alias, target in _ALIASES.items():
        sys.modules[alias] = importlib.import_module(target)
