"""BiEncoder / MultiTextBiEncoder — host-side mirrors of reference models/audio_text_model.py:16-98 and
:101-229 (orchestration only: audio encoder -> text encoder -> match function [-> clip-level pooling])."""
from __future__ import annotations

import sys
from typing import List, Optional

import torch.nn as nn

from .base import LoadPretrainedMixin
from .utils import pool_with_lens


class BiEncoder(nn.Module, LoadPretrainedMixin):
    def __init__(self, audio_encoder: nn.Module, text_encoder: nn.Module, match_fn: nn.Module,
                 shared_dim: int, cross_encoder: Optional[nn.Module] = None, add_proj: bool = False,
                 upsample: bool = False, freeze_audio_encoder: bool = False,
                 freeze_text_encoder: bool = False, pretrained: Optional[str] = None):
        super().__init__()
        self.audio_encoder = audio_encoder
        self.text_encoder = text_encoder
        self.match_fn = match_fn
        self.cross_encoder = cross_encoder
        if cross_encoder is not None:
            raise NotImplementedError("cross_encoder is outside the cnn8rnn-w2vmean hot path (SURVEY.md §8f)")
        if audio_encoder.embed_dim != text_encoder.embed_dim or add_proj:
            raise NotImplementedError("audio_proj/text_proj (add_proj or mismatched embed dims) are "
                                      "outside the cnn8rnn-w2vmean hot path (SURVEY.md §8f)")
        if upsample:
            raise NotImplementedError("upsample=True is outside the cnn8rnn-w2vmean hot path")
        self.interpolate_ratio = self.audio_encoder.downsample_ratio
        self.upsample = upsample
        self.shared_dim = shared_dim
        if pretrained is not None and type(self) is BiEncoder:
            self.load_pretrained(pretrained)
        if freeze_audio_encoder:
            for param in self.audio_encoder.parameters():
                param.requires_grad = False
        if freeze_text_encoder:
            for param in self.text_encoder.parameters():
                param.requires_grad = False

    def forward(self, input_dict):
        """keys in input_dict: waveform, waveform_len, text, text_len, specaug"""
        audio_output = self.audio_encoder(input_dict)
        audio_emb = audio_output["embedding"]
        text_emb = self.text_encoder(input_dict)
        forward_dict = {"audio_emb": audio_emb, "text_emb": text_emb,
                        "audio_len": audio_output["length"]}
        if "text_len" in input_dict:
            forward_dict["text_len"] = input_dict["text_len"]
        frame_sim = self.match_fn(forward_dict)      # [batch_size, max_len]
        length = audio_output["length"]
        return {"frame_sim": frame_sim, "length": length}


class MultiTextBiEncoder(BiEncoder):
    """Every clip against ``text_num`` phrases (the weakly supervised runners, reference
    python_scripts/training/run_weak_phrase.py:39-60).  Same constructor and dict contract as the reference
    (models/audio_text_model.py:101-229): ``input_dict[text_forward_keys[0]]`` is [B, text_num, N]; returns
    ``frame_sim`` [B, T', text_num], ``clip_sim`` [B, text_num] and ``length``.  The audio embedding is not
    expanded per phrase and ``safe_size`` chunking is therefore unnecessary (accepted and ignored)."""

    def __init__(self, audio_encoder: nn.Module, text_encoder: nn.Module, match_fn: nn.Module, shared_dim: int,
                 text_forward_keys: "List[str]", cross_encoder: "nn.Module | None" = None,
                 pooling: str = "linear_softmax", add_proj: bool = False, upsample: bool = False,
                 freeze_audio_encoder: bool = False, freeze_text_encoder: bool = False,
                 safe_size: "int | None" = None, pretrained: "str | None" = None,
                 output_fn: callable = sys.stdout.write):
        super().__init__(audio_encoder=audio_encoder, text_encoder=text_encoder, match_fn=match_fn,
                         shared_dim=shared_dim, cross_encoder=cross_encoder, add_proj=add_proj, upsample=upsample,
                         freeze_audio_encoder=freeze_audio_encoder, freeze_text_encoder=freeze_text_encoder)
        self.text_forward_keys = text_forward_keys
        if "text_len" not in text_forward_keys:
            self.text_forward_keys.append("text_len")
        self.pooling = pooling
        self.safe_size = safe_size
        if pretrained is not None and type(self) is MultiTextBiEncoder:
            self.load_pretrained(pretrained, output_fn)

    def process_state_dict(self, model_dict, pretrained_dict, output_fn, model_name):
        pretrained_dict = pretrained_dict["model"]
        return super().process_state_dict(model_dict, pretrained_dict, output_fn, model_name)

    def forward(self, input_dict):
        import torch
        audio_output = self.audio_encoder(input_dict)
        audio_emb = audio_output["embedding"]                        # [B, T', D]
        batch_size = audio_emb.size(0)
        text_num = input_dict[self.text_forward_keys[0]].shape[1]
        text_forward_dict = {}
        for key in self.text_forward_keys:
            x = torch.as_tensor(input_dict[key])
            text_forward_dict[key] = x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])
        text_emb = self.text_encoder(text_forward_dict)
        seq_emb = text_emb["seq_emb"].view(batch_size, text_num, -1)
        if not hasattr(self.match_fn, "forward_multi"):
            raise NotImplementedError("MultiTextBiEncoder (B200) needs a match function with forward_multi "
                                      "(models.match.DotProduct)")
        frame_sim = self.match_fn.forward_multi(audio_emb, seq_emb)  # [B, T', text_num]
        length = audio_output["length"]
        clip_sim = pool_with_lens(frame_sim, length, self.pooling)
        return {"frame_sim": frame_sim, "clip_sim": clip_sim, "length": length}
