"""Inference façade with the call signature of the reference's Hugging Face model
(models/hf_modeling_grounding.py:319-352, README.md:8-39): ``model(audio, audio_len, text) -> frame_sim``.

The in-repo reference façade wires the CLAP text tower (not built here); this is the same
surface for the cnn8rnn-w2vmean model of the hot path: Cnn8Rnn + EmbeddingAgg(mean) + DotProduct with the
reference's DictTokenizer (whitespace tokens, ``<unk>`` for unknown words).  ``config`` carries the reference's
``sample_rate`` / ``shared_dim`` plus the vocabulary."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Union

import torch
import torch.nn as nn

from ..datasets.text_tokenizer import DictTokenizer
from .audio_encoder import Cnn8Rnn
from .audio_text_model import BiEncoder
from .match import DotProduct
from .text_encoder import EmbeddingAgg


@dataclass
class Cnn8RnnW2vMeanGroundingConfig:
    sample_rate: int = 32000
    shared_dim: int = 512
    vocabulary: Union[str, Dict[str, int]] = field(default_factory=lambda: {"<pad>": 0, "<unk>": 1})
    embed_dim: int = 512


class Cnn8RnnW2vMeanGroundingModel(nn.Module):
    config_class = Cnn8RnnW2vMeanGroundingConfig

    def __init__(self, config: Cnn8RnnW2vMeanGroundingConfig):
        super().__init__()
        self.config = config
        self.text_tokenizer = DictTokenizer(config.vocabulary)
        self.model = BiEncoder(
            audio_encoder=Cnn8Rnn(sample_rate=config.sample_rate),
            text_encoder=EmbeddingAgg(len(self.text_tokenizer.vocabulary), config.embed_dim),
            match_fn=DotProduct(),
            shared_dim=config.shared_dim)

    @property
    def device(self):
        return next(self.parameters()).device

    def forward(self, audio: torch.Tensor, audio_len, text: List[str]) -> torch.Tensor:
        device = self.device
        tokens = self.text_tokenizer(text)
        input_dict = {
            "waveform": audio.to(device),
            "waveform_len": audio_len,
            "specaug": False,
            "text": tokens["text"].to(device),
            "text_len": tokens["text_len"].to(device),
        }
        return self.model(input_dict)["frame_sim"]
