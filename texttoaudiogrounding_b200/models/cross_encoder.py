"""Cross encoder — host-side mirror of reference models/cross_encoder.py:5-79 (Seq2SeqAttention, CrossGating,
CrossAttentionGating): every audio frame attends over the text tokens with additive (tanh) attention, then audio
and attended text gate each other.

The reference builds q_repeat / kv_repeat / cat([B, T*N, 2E]) and tanh(h2attn(.)) ([B, T*N, E]) in HBM.  h2attn is
linear, so h2attn(cat(q, kv)) = q W_q^T + (kv W_k^T + b) with W = [W_q | W_k]: two small GEMMs ([B*T, E] and
[B*N, E]) and one kernel (csrc/attn.cu, additive_attn_*) that forms tanh / v-dot / masks / softmax / weighted sum per
frame in registers.  Same constructors, parameters and state-dict keys as the reference."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import nn_ops
from .utils import lens_to_device


class Seq2SeqAttention(nn.Module):
    def __init__(self, d_q, d_kv, d_attn):
        super().__init__()
        self.h2attn = nn.Linear(d_q + d_kv, d_attn)
        self.v = nn.Parameter(torch.randn(d_attn))
        self.d_q = d_q

    def forward(self, query, kv, query_len, kv_len):
        if not query.is_cuda:
            raise RuntimeError("Seq2SeqAttention (B200) needs CUDA tensors: there is no CPU fallback")
        dev = query.device
        W = self.h2attn.weight
        hq = nn_ops.linear(query, W[:, :self.d_q].contiguous(), None)               # [B, T, d_attn]
        hk = nn_ops.linear(kv, W[:, self.d_q:].contiguous(), self.h2attn.bias)      # [B, N, d_attn]
        return nn_ops.additive_attention(hq, hk, self.v, kv, lens_to_device(query_len, dev).contiguous(),
                                         lens_to_device(kv_len, dev).contiguous())


class CrossGating(nn.Module):
    def __init__(self, d_model) -> None:
        super().__init__()
        self.fc_u = nn.Linear(d_model, d_model)
        self.fc_s = nn.Linear(d_model, d_model)

    def forward(self, u, s):
        s_out = nn_ops.sigmoid_gate(s, nn_ops.linear(u, self.fc_u.weight, self.fc_u.bias))
        u_out = nn_ops.sigmoid_gate(u, nn_ops.linear(s, self.fc_s.weight, self.fc_s.bias))
        return u_out, s_out


class CrossAttentionGating(nn.Module):
    def __init__(self, embed_dim):
        super().__init__()
        self.attn = Seq2SeqAttention(embed_dim, embed_dim, embed_dim)
        self.gating = CrossGating(embed_dim)

    def forward(self, input_dict):
        audio_emb = input_dict["audio_emb"]
        text_emb = input_dict["text_emb"]
        audio_len = input_dict["audio_len"]
        text_len = input_dict["text_len"]
        if isinstance(text_emb, dict):
            text_emb = text_emb["token_emb"]
        text_emb = self.attn(audio_emb, text_emb, audio_len, text_len)
        audio_emb, text_emb = self.gating(audio_emb, text_emb)
        return {"audio_emb": audio_emb, "text_emb": {"token_emb": text_emb}}
