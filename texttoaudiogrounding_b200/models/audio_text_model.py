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
        if audio_encoder.embed_dim != text_encoder.embed_dim or add_proj:
            # parameter containers; applied with nn_ops.linear (models/audio_text_model.py:35-46, 82-97)
            self.audio_proj = nn.Linear(audio_encoder.embed_dim, shared_dim)
            self.text_proj = nn.Linear(text_encoder.embed_dim, shared_dim)
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
        if self.cross_encoder is not None:
            forward_dict.update(self.cross_encoder(forward_dict))     # audio_emb, text_emb (per-frame tokens)
        if hasattr(self, "audio_proj"):
            from . import nn_ops
            forward_dict["audio_emb"] = nn_ops.linear(forward_dict["audio_emb"], self.audio_proj.weight,
                                                      self.audio_proj.bias)
            text_emb = forward_dict["text_emb"]
            for key in ("seq_emb", "token_emb"):
                if key in text_emb:
                    text_emb[key] = nn_ops.linear(text_emb[key], self.text_proj.weight, self.text_proj.bias)
        frame_sim = self.match_fn(forward_dict)      # [batch_size, max_len]
        length = audio_output["length"]
        if self.interpolate_ratio != 1 and self.upsample:
            from . import nn_ops
            frame_sim = nn_ops.upsample_linear(frame_sim, frame_sim.size(1) * self.interpolate_ratio)
            length = length * self.interpolate_ratio
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
        if "model" in pretrained_dict:          # the mixin may already have unwrapped the checkpoint
            pretrained_dict = pretrained_dict["model"]
        return super().process_state_dict(model_dict, pretrained_dict, output_fn, model_name)

    def _text_proj(self, text_emb):
        if hasattr(self, "text_proj"):
            from . import nn_ops
            for key in ("seq_emb", "token_emb"):
                if key in text_emb:
                    text_emb[key] = nn_ops.linear(text_emb[key], self.text_proj.weight, self.text_proj.bias)

    def forward(self, input_dict):
        """Order of operations as in the reference (models/audio_text_model.py:147-229): audio_proj on the audio
        embedding FIRST, text encoder on the flattened [B*text_num, ...] phrases, cross encoder on the per-pair
        batch, text_proj, match function, clip-level pooling, optional upsampling."""
        import torch
        from . import nn_ops
        audio_output = self.audio_encoder(input_dict)
        audio_emb = audio_output["embedding"]                        # [B, T', D]
        if hasattr(self, "audio_proj"):
            audio_emb = nn_ops.linear(audio_emb, self.audio_proj.weight, self.audio_proj.bias)
        batch_size = audio_emb.size(0)
        text_num = input_dict[self.text_forward_keys[0]].shape[1]
        text_forward_dict = {}
        for key in self.text_forward_keys:
            x = torch.as_tensor(input_dict[key])
            text_forward_dict[key] = x.reshape(x.shape[0] * x.shape[1], *x.shape[2:])
        text_emb = self.text_encoder(text_forward_dict)
        length = audio_output["length"]
        fast = (self.cross_encoder is None and hasattr(self.match_fn, "forward_multi")
                and getattr(self.match_fn, "text_level", None) == "seq" and not getattr(self.match_fn, "l2norm", True))
        if fast:
            # every clip against its phrases with the audio embedding read in place (no [B*n, T', D] expansion)
            self._text_proj(text_emb)
            seq_emb = text_emb["seq_emb"].view(batch_size, text_num, -1)
            frame_sim = self.match_fn.forward_multi(audio_emb, seq_emb)       # [B, T', text_num]
        else:
            # general route: one (clip, phrase) pair per row, exactly the reference's expansion
            audio_rep = audio_emb.unsqueeze(1).expand(-1, text_num, -1, -1).reshape(-1, *audio_emb.shape[1:])
            audio_len = torch.as_tensor(length).repeat_interleave(text_num)
            forward_dict = {"audio_emb": audio_rep, "text_emb": text_emb, "audio_len": audio_len,
                            "text_len": text_forward_dict["text_len"]}
            if self.cross_encoder is not None:
                forward_dict.update(self.cross_encoder(forward_dict))
            self._text_proj(forward_dict["text_emb"])
            frame_sim = self.match_fn(forward_dict)                           # [B * text_num, T']
            frame_sim = frame_sim.reshape(batch_size, text_num, -1).transpose(1, 2)
        clip_sim = pool_with_lens(frame_sim, length, self.pooling)
        if self.interpolate_ratio != 1 and self.upsample:
            # the reference sizes the interpolation by frame_sim.size(-1) of the [B, T', text_num] tensor, i.e. by
            # the NUMBER OF PHRASES (models/audio_text_model.py:217-219): T' frames are resampled to
            # text_num * ratio frames.  Reproduced as is — results must equal the reference's.
            frame_sim = nn_ops.upsample_linear(frame_sim, frame_sim.size(-1) * self.interpolate_ratio)
            length = length * self.interpolate_ratio
        return {"frame_sim": frame_sim, "clip_sim": clip_sim, "length": length}


class _AudioTextAlign(nn.Module):
    """Shared constructor of the sentence-level alignment models (reference models/audio_text_model.py:843-869,
    907-935): all-pairs similarity ``match_fn`` (models.align.DotProduct) + a ``sim_pooling`` module."""

    def __init__(self, audio_encoder: nn.Module, text_encoder: nn.Module, match_fn: nn.Module, sim_pooling: nn.Module,
                 shared_dim: int, cross_encoder: Optional[nn.Module] = None, add_proj: bool = False,
                 freeze_audio_encoder: bool = False, freeze_text_encoder: bool = False):
        super().__init__()
        self.audio_encoder = audio_encoder
        self.text_encoder = text_encoder
        self.match_fn = match_fn
        self.sim_pooling = sim_pooling
        self.shared_dim = shared_dim
        if cross_encoder is not None:
            raise NotImplementedError("cross_encoder is not on the B200 path (SURVEY.md §8f rank 2)")
        if audio_encoder.embed_dim != text_encoder.embed_dim or add_proj:
            raise NotImplementedError("audio_proj/text_proj (add_proj or mismatched embed dims) are not on the "
                                      "B200 path (SURVEY.md §8f rank 2)")
        if freeze_audio_encoder:
            for param in self.audio_encoder.parameters():
                param.requires_grad = False
        if freeze_text_encoder:
            for param in self.text_encoder.parameters():
                param.requires_grad = False

    def _finish(self, input_dict, sim_matrix, audio_len, text_len):
        sim = self.sim_pooling({"sim": sim_matrix, "audio_len": audio_len, "text_len": text_len})
        output = {"sim": sim}
        if input_dict.get("output_matrix", False):
            output["sim_matrix"] = sim_matrix.materialize() if hasattr(sim_matrix, "materialize") else sim_matrix
        return output


class AudioTextAlignByWord(_AudioTextAlign):
    """Every clip against every caption, word by word (reference models/audio_text_model.py:843-904)."""

    def __init__(self, audio_encoder, text_encoder, match_fn, sim_pooling, shared_dim, add_proj=False,
                 freeze_audio_encoder=False, freeze_text_encoder=False):
        super().__init__(audio_encoder, text_encoder, match_fn, sim_pooling, shared_dim, None, add_proj,
                         freeze_audio_encoder, freeze_text_encoder)

    def forward(self, input_dict):
        audio_output = self.audio_encoder(input_dict)
        audio_emb = audio_output["embedding"]                     # [bs, n_seg, emb_dim]
        word_emb = self.text_encoder(input_dict)["token_emb"]     # [bs, n_word, emb_dim]
        sim_matrix = self.match_fn(audio_emb, word_emb)           # [bs, bs, n_seg, n_word] (deferred)
        return self._finish(input_dict, sim_matrix, audio_output["length"], input_dict["text_len"])


class AudioTextAlignByPhrase(_AudioTextAlign):
    """Every clip against the phrases of every caption (reference models/audio_text_model.py:907-976):
    ``input_dict[text_key]`` holds all phrases of the batch [txt_num, max_txt_len], ``{text_key}_num`` how many
    belong to each clip; the phrase embeddings are regrouped to [bs, max_txt_num, emb_dim] (zero padded)."""

    def forward(self, input_dict):
        import torch
        audio_output = self.audio_encoder(input_dict)
        audio_emb = audio_output["embedding"]
        text_key = input_dict["text_key"]
        phrases_emb = self.text_encoder({"text": input_dict[text_key], "text_len": input_dict[f"{text_key}_len"]})
        phrases_num = input_dict[f"{text_key}_num"]
        if isinstance(phrases_num, torch.Tensor):
            phrases_num = phrases_num.tolist()
        seq_emb = torch.split(phrases_emb["seq_emb"], [int(n) for n in phrases_num], dim=0)
        seq_emb = nn.utils.rnn.pad_sequence(seq_emb, batch_first=True)      # [bs, max_txt_num, emb_dim]
        sim_matrix = self.match_fn(audio_emb, seq_emb)
        return self._finish(input_dict, sim_matrix, audio_output["length"], phrases_num)
