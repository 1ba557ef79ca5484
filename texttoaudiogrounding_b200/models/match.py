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
