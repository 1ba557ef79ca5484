"""LoadPretrainedMixin — same semantics as reference models/base.py:9-47 and
utils/train_util.py:219-297: load a checkpoint, keep the entries whose key AND shape match,
skip the rest silently."""
from pathlib import Path
from typing import Callable
import sys

import torch


def merge_matched_keys(model_dict, pretrained_dict, output_fn: Callable = sys.stdout.write,
                       model_name: str = ""):
    matched, mismatched = {}, []
    for key, value in pretrained_dict.items():
        if key in model_dict and model_dict[key].shape == value.shape:
            matched[key] = value
        else:
            mismatched.append(key)
    unmatched = [k for k in model_dict if k not in matched]
    if output_fn is not None and (mismatched or unmatched):
        output_fn(f"[{model_name}] loaded {len(matched)} tensors; "
                  f"{len(mismatched)} checkpoint keys skipped; {len(unmatched)} model keys kept\n")
    state = dict(model_dict)
    state.update(matched)
    return state


class LoadPretrainedMixin:
    def process_state_dict(self, model_dict, pretrained_dict, output_fn, model_name):
        return merge_matched_keys(model_dict, pretrained_dict, output_fn, model_name)

    def load_pretrained(self, ckpt_path: "str | Path", output_fn: Callable = sys.stdout.write):
        pretrained = torch.load(ckpt_path, map_location="cpu")
        if isinstance(pretrained, dict) and "model" in pretrained and \
                not any(k in self.state_dict() for k in pretrained):
            pretrained = pretrained["model"]
        state = self.process_state_dict(self.state_dict(), pretrained, output_fn,
                                        self.__class__.__name__)
        self.load_state_dict(state)
