"""Frame-wise audio<->text match — host-side mirror of reference models/match.py:36-60
(DotProduct).  score = sigmoid(<audio[b,t,:], text[b,:]> / sqrt(D)).clamp(1e-7, 1) runs as a
warp-shuffle reduction kernel (csrc/head.cu), with an analytic backward."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..ops import call
from . import nn_ops
from .utils import lens_to_device


class _DotSigmoidFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, audio, seq, scale):
        B, T, D = audio.shape
        sim = torch.empty(B, T, device=audio.device, dtype=torch.float32)
        call("tag_dot_sigmoid_fwd", audio, seq, sim, None, B, T, D, scale)
        ctx.save_for_backward(audio, seq, sim)
        ctx.scale = scale
        return sim

    @staticmethod
    def backward(ctx, d_sim):
        audio, seq, sim = ctx.saved_tensors
        B, T, D = audio.shape
        d_audio = torch.empty_like(audio)
        d_seq = torch.empty_like(seq)
        ws = torch.empty(B, T, device=audio.device, dtype=torch.float32)
        call("tag_dot_sigmoid_bwd", d_sim.contiguous(), sim, audio, seq, d_audio, d_seq, ws, B, T, D,
             ctx.scale)
        return d_audio, d_seq, None


class _MultiDotSigmoidFunction(torch.autograd.Function):
    """sim[b,t,j] for every (clip, phrase) pair; the audio embedding is read in place (the reference expands it
    to [B*n, T, D], models/audio_text_model.py:166-170)."""

    @staticmethod
    def forward(ctx, audio, seq, scale):
        B, T, D = audio.shape
        n = seq.shape[1]
        sim = torch.empty(B, T, n, device=audio.device, dtype=torch.float32)
        call("tag_multi_dot_sigmoid_fwd", audio, seq, sim, B, T, n, D, scale)
        ctx.save_for_backward(audio, seq, sim)
        ctx.scale = scale
        return sim

    @staticmethod
    def backward(ctx, d_sim):
        audio, seq, sim = ctx.saved_tensors
        B, T, D = audio.shape
        n = seq.shape[1]
        d_audio = torch.empty_like(audio)
        d_seq = torch.empty_like(seq)
        ws = torch.empty_like(sim)
        call("tag_multi_dot_sigmoid_bwd", d_sim.contiguous(), sim, audio, seq, d_audio, d_seq, ws, B, T, n, D,
             ctx.scale)
        return d_audio, d_seq, None


class _MatchNormFunction(torch.autograd.Function):
    """mode 1: cosine -> scale -> sigmoid -> clamp; mode 2 / 3: exp(-|a^ - s^|) with / without L2 normalisation."""

    @staticmethod
    def forward(ctx, audio, seq, mode, scale):
        B, T, D = audio.shape
        sim = torch.empty(B, T, device=audio.device, dtype=torch.float32)
        call("tag_match_norm_fwd", audio, seq, sim, B, T, D, mode, scale)
        ctx.save_for_backward(audio, seq, sim)
        ctx.cfg = (mode, scale)
        return sim

    @staticmethod
    def backward(ctx, d_sim):
        audio, seq, sim = ctx.saved_tensors
        mode, scale = ctx.cfg
        B, T, D = audio.shape
        d_audio = torch.empty_like(audio)
        d_seq = torch.zeros_like(seq)
        call("tag_match_norm_bwd", d_sim.contiguous(), sim, audio, seq, d_audio, d_seq, B, T, D, mode, scale)
        return d_audio, d_seq, None, None


MULTI_MAX_PHRASES = 64      # phrases per clip handled by one kernel launch


class DotProduct(nn.Module):
    def __init__(self, l2norm=False, scale=True, text_level="seq") -> None:
        super().__init__()
        self.l2norm = l2norm
        self.scale = scale
        self.text_level = text_level

    def forward(self, input_dict):
        audio = input_dict["audio_emb"]  # [bs, n_seg, dim]
        text = input_dict["text_emb"]
        if not audio.is_cuda:
            raise RuntimeError("DotProduct (B200) needs CUDA tensors: there is no CPU fallback")
        scale = 1.0 / math.sqrt(audio.size(-1)) if self.scale else 1.0
        if self.l2norm:
            # cosine similarity (F.normalize on both sides, models/match.py:51-53)
            if self.text_level != "seq":
                raise NotImplementedError("DotProduct(l2norm=True) is built for text_level='seq'")
            return _MatchNormFunction.apply(audio.float().contiguous(), text["seq_emb"].float().contiguous(), 1, scale)
        if self.text_level == "seq":
            text = text["seq_emb"]      # [bs, dim]
        elif self.text_level == "token":
            # per-frame text embeddings from a cross encoder: [bs, n_seg, dim] against [bs, n_seg, dim]
            text = text["token_emb"]
            if text.shape != audio.shape:
                raise RuntimeError(f"The size of tensor a {tuple(audio.shape)} must match the size of tensor b "
                                   f"{tuple(text.shape)}")
            return nn_ops.rowdot_sigmoid(audio, text, scale)
        else:
            raise KeyError(self.text_level)
        return _DotSigmoidFunction.apply(audio.float().contiguous(), text.float().contiguous(), scale)

    def forward_multi(self, audio: torch.Tensor, seq: torch.Tensor) -> torch.Tensor:
        """audio [B, T, D], seq [B, n, D] -> frame_sim [B, T, n] (every clip against its n phrases)."""
        if self.text_level != "seq" or self.l2norm:
            raise NotImplementedError("only text_level='seq', l2norm=False is on the B200 path")
        if not audio.is_cuda:
            raise RuntimeError("DotProduct (B200) needs CUDA tensors: there is no CPU fallback")
        scale = 1.0 / math.sqrt(audio.size(-1)) if self.scale else 1.0
        audio = audio.float().contiguous()
        seq = seq.float().contiguous()
        chunks = [_MultiDotSigmoidFunction.apply(audio, seq[:, i:i + MULTI_MAX_PHRASES].contiguous(), scale)
                  for i in range(0, seq.shape[1], MULTI_MAX_PHRASES)]
        return chunks[0] if len(chunks) == 1 else torch.cat(chunks, dim=2)


class CrossAttention(nn.Module):
    """Cross-attention match head — mirror of reference models/match.py:63-88: every audio frame attends over the
    text tokens (nn.MultiheadAttention with key padding mask), then residual + dropout -> LayerNorm -> Linear(E, 1)
    -> sigmoid.  Same constructor and state-dict keys; the modules hold the parameters only."""

    def __init__(self, embed_dim, num_heads, dropout, kvdim=None) -> None:
        super().__init__()
        if kvdim is not None and kvdim != embed_dim:
            raise NotImplementedError("CrossAttention (B200): kvdim must equal embed_dim (packed in_proj)")
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout, batch_first=True, kdim=kvdim, vdim=kvdim)
        self.dropout = nn.Dropout(dropout)
        self.norm = nn.LayerNorm(embed_dim)
        self.linear = nn.Linear(embed_dim, 1)

    def forward(self, input_dict):
        audio = input_dict["audio_emb"]  # [bs, n_seg, dim]
        text = input_dict["text_emb"]["token_emb"]
        if not audio.is_cuda:
            raise RuntimeError("CrossAttention (B200) needs CUDA tensors: there is no CPU fallback")
        text_len = lens_to_device(input_dict["text_len"], audio.device).contiguous()
        audio = audio.float().contiguous()
        out = nn_ops.multi_head_attention(self.attn, audio, text.float().contiguous(), text.float().contiguous(),
                                          text_len, self.training)
        p = self.dropout.p if self.training else 0.0
        return nn_ops.ln_linear_sigmoid(audio, out, self.norm, self.linear, p)


class ExpNegL2(nn.Module):
    """exp(-|a - s|) on (optionally L2-normalised) embeddings — mirror of reference models/match.py:10-33
    (the match function of eg_configs/strongly_supervised/audiogrounding/biencoder/cdur_w2vmean.yaml)."""

    def __init__(self, l2norm=True, text_level="seq") -> None:
        super().__init__()
        self.l2norm = l2norm
        self.text_level = text_level

    def forward(self, input_dict):
        audio = input_dict["audio_emb"]
        if self.text_level != "seq":
            raise NotImplementedError("ExpNegL2 (B200) is built for text_level='seq'")
        if not audio.is_cuda:
            raise RuntimeError("ExpNegL2 (B200) needs CUDA tensors: there is no CPU fallback")
        text = input_dict["text_emb"]["seq_emb"]
        return _MatchNormFunction.apply(audio.float().contiguous(), text.float().contiguous(),
                                        2 if self.l2norm else 3, 1.0)
