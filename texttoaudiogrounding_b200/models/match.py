"""Frame-wise audio<->text match — host-side mirror of reference models/match.py:36-60
(DotProduct).  score = sigmoid(<audio[b,t,:], text[b,:]> / sqrt(D)).clamp(1e-7, 1) runs as a
warp-shuffle reduction kernel (csrc/head.cu), with an analytic backward."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ..ops import call


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
        if self.text_level == "seq":
            text = text["seq_emb"]      # [bs, dim]
        else:
            raise NotImplementedError("text_level='token' is outside the cnn8rnn-w2vmean hot path")
        if self.l2norm:
            raise NotImplementedError("l2norm=True is outside the cnn8rnn-w2vmean hot path")
        if not audio.is_cuda:
            raise RuntimeError("DotProduct (B200) needs CUDA tensors: there is no CPU fallback")
        scale = 1.0 / math.sqrt(audio.size(-1)) if self.scale else 1.0
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
