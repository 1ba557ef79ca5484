"""Inference façade with the call signature of the reference's Hugging Face model
(models/hf_modeling_grounding.py:319-352, README.md:8-39): ``model(audio, audio_len, text) -> frame_sim``.

The in-repo reference façade wires the CLAP text tower (Cnn8RnnLaionClapGroundingModel below); this is the same
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


try:                                    # transformers is optional for everything but the CLAP graph
    from transformers import PretrainedConfig, PreTrainedModel
except Exception:                       # pragma: no cover
    PretrainedConfig, PreTrainedModel = object, nn.Module


class Cnn8RnnLaionClapGroundingConfig(PretrainedConfig):
    """reference models/hf_modeling_grounding.py:305-316 (a ``PretrainedConfig`` with the same fields, so a released
    ``config.json`` loads).  Offline extension: ``text_config`` (a ``ClapTextConfig`` or its dict) builds a randomly
    initialised text tower instead of downloading ``text_encoder_name``; ``text_encoder_name`` may also be given as a
    ``ClapConfig`` / ``ClapTextConfig`` object directly."""
    model_type = "cnn8rnn_laionclap_grounding"

    def __init__(self, sample_rate: int = 32000, shared_dim: int = 512,
                 text_encoder_name="laion/clap-htsat-fused", text_config=None, **kwargs):
        if not isinstance(text_encoder_name, str) and text_config is None:
            text_config, text_encoder_name = text_encoder_name, "laion/clap-htsat-fused"
        if text_config is not None and not isinstance(text_config, dict):
            text_config = getattr(text_config, "text_config", text_config).to_dict()
        self.sample_rate = sample_rate
        self.shared_dim = shared_dim
        self.text_encoder_name = text_encoder_name
        self.text_config = text_config
        super().__init__(**kwargs)


class Cnn8RnnLaionClapGroundingModel(PreTrainedModel):
    """``model(audio, audio_len, text) -> frame_sim`` of the released CLAP checkpoints (reference
    models/hf_modeling_grounding.py:319-352): Cnn8Rnn + LaionClapEncoder + DotProduct inside
    BiEncoder(add_proj=True); a ``PreTrainedModel``, so ``save_pretrained`` / ``from_pretrained`` directories
    interchange with the reference's (identical parameter names under ``model.``).  ``text`` is a list of strings
    (needs the Hugging Face tokenizer of ``text_encoder_name``) or an already tokenised dict with ``input_ids`` /
    ``attention_mask``."""
    config_class = Cnn8RnnLaionClapGroundingConfig
    base_model_prefix = "model"

    def __init__(self, config: Cnn8RnnLaionClapGroundingConfig, text_tokenizer=None):
        super().__init__(config)
        from .text_encoder import LaionClapEncoder
        self.text_tokenizer = text_tokenizer
        if config.text_config is not None:
            from transformers import ClapTextConfig
            tower = ClapTextConfig(**{k: v for k, v in config.text_config.items() if k != "model_type"})
        else:
            tower = config.text_encoder_name
            if text_tokenizer is None:
                from transformers import AutoTokenizer
                self.text_tokenizer = AutoTokenizer.from_pretrained(config.text_encoder_name)
        self.model = BiEncoder(
            audio_encoder=Cnn8Rnn(sample_rate=config.sample_rate),
            text_encoder=LaionClapEncoder(model_type=tower),
            match_fn=DotProduct(),
            shared_dim=config.shared_dim,
            add_proj=True)
        if hasattr(self, "post_init"):
            self.post_init()

    def _init_weights(self, module):       # the sub-modules initialise themselves (reference init distributions)
        pass

    def forward(self, audio: torch.Tensor, audio_len, text) -> torch.Tensor:
        device = next(self.parameters()).device
        if isinstance(text, dict):
            tokens = {k: torch.as_tensor(v).to(device) for k, v in text.items()}
        else:
            if self.text_tokenizer is None:
                raise RuntimeError("no tokenizer: pass text_tokenizer=... or a tokenised dict")
            tokens = dict(self.text_tokenizer(text, padding=True, return_tensors="pt", truncation=True).to(device))
        tokens["text_len"] = tokens["attention_mask"].sum(dim=-1)
        input_dict = {"waveform": audio.to(device), "waveform_len": audio_len, "specaug": False}
        input_dict.update(tokens)
        return self.model(input_dict)["frame_sim"]


def register_auto_classes() -> None:
    """``AutoConfig`` / ``AutoModel`` resolve ``model_type: cnn8rnn_laionclap_grounding`` to the B200 classes, so the
    reference README's ``AutoModel.from_pretrained(path)`` call (README.md:8-39) works on a local checkpoint
    directory without ``trust_remote_code``."""
    from transformers import AutoConfig, AutoModel
    try:
        AutoConfig.register(Cnn8RnnLaionClapGroundingConfig.model_type, Cnn8RnnLaionClapGroundingConfig)
    except ValueError:
        pass                                   # already registered
    try:
        AutoModel.register(Cnn8RnnLaionClapGroundingConfig, Cnn8RnnLaionClapGroundingModel)
    except ValueError:
        pass
