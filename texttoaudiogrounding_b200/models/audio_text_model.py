"""BiEncoder — host-side mirror of reference models/audio_text_model.py:16-98 (orchestration
only: audio encoder -> text encoder -> match function)."""
from __future__ import annotations

from typing import Optional

import torch.nn as nn

from .base import LoadPretrainedMixin


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
