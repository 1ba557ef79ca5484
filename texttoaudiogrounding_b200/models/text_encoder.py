"""Word-embedding text encoder — host-side mirror of reference models/text_encoder.py:14-43
(EmbeddingLayer) and :61-88 (EmbeddingAgg, aggregation="mean").  The gather + masked mean and
its scatter-add backward run in csrc/head.cu."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

from ..ops import call
from .utils import init_weights, lens_to_device


class _EmbedMeanFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weight, tokens, lens):
        B, N = tokens.shape
        V, D = weight.shape
        token_emb = torch.empty(B, N, D, device=weight.device, dtype=torch.float32)
        seq_emb = torch.empty(B, D, device=weight.device, dtype=torch.float32)
        call("tag_embed_mean_fwd", tokens, lens, weight, token_emb, seq_emb, B, N, D, V)
        ctx.save_for_backward(tokens, lens)
        ctx.shape = (V, D)
        ctx.set_materialize_grads(False)        # an unused output arrives as None, not as a zero tensor
        return token_emb, seq_emb

    @staticmethod
    def backward(ctx, d_token, d_seq):
        tokens, lens = ctx.saved_tensors
        V, D = ctx.shape
        B, N = tokens.shape
        d_w = torch.zeros(V, D, device=tokens.device, dtype=torch.float32)
        if d_seq is not None:
            call("tag_embed_mean_bwd", tokens, lens, d_seq.contiguous(), d_w, B, N, D, V)
        if d_token is not None:          # word-level consumers (AudioTextAlignByWord): nn.Embedding backward
            call("tag_embed_token_bwd", tokens, d_token.contiguous(), d_w, B, N, D, V)
        return d_w, None, None


class EmbeddingLayer(nn.Module):
    def __init__(self, vocab_size: int, embed_dim: int, pretrained_embedding: str = None,
                 freeze_embedding: bool = False):
        super().__init__()
        self.embed_dim = embed_dim
        self.core = nn.Embedding(vocab_size, embed_dim)
        self.apply(init_weights)
        if pretrained_embedding is not None:
            self.load_pretrained_embedding(pretrained_embedding, freeze_embedding)

    def load_pretrained_embedding(self, weight: str, freeze: bool = True):
        weight = np.load(weight)
        assert weight.shape == self.core.weight.shape, \
            f"expect embedding with shape {self.core.weight.shape} " \
            f"but {weight.shape} is given"
        weight = torch.as_tensor(weight, dtype=torch.float)
        self.core = nn.Embedding.from_pretrained(weight, freeze)

    def forward(self, input_dict: Dict):
        tokens = input_dict["text"].long()
        lens = torch.full((tokens.shape[0],), tokens.shape[1], device=tokens.device, dtype=torch.long)
        return _EmbedMeanFunction.apply(self.core.weight, tokens.contiguous(), lens)[0]


class EmbeddingAgg(nn.Module):
    def __init__(self, vocab_size, embed_dim, pretrained_embedding: str = None,
                 freeze_embedding: bool = False, aggregation: str = "mean"):
        super().__init__()
        self.embedding = EmbeddingLayer(vocab_size, embed_dim, pretrained_embedding, freeze_embedding)
        self.embed_dim = self.embedding.embed_dim
        self.agg = aggregation
        if aggregation == "attention":
            raise NotImplementedError("aggregation='attention' is outside the cnn8rnn-w2vmean hot path")

    def forward(self, input_dict):
        if self.agg != "mean":
            raise Exception(f"{self.agg} not supported")
        weight = self.embedding.core.weight
        if not weight.is_cuda:
            raise RuntimeError("EmbeddingAgg (B200) needs CUDA tensors: there is no CPU fallback")
        tokens = input_dict["text"].long().to(weight.device).contiguous()
        lens = lens_to_device(input_dict["text_len"], weight.device).contiguous()
        token_emb, seq_emb = _EmbedMeanFunction.apply(weight, tokens, lens)
        return {"token_emb": token_emb, "seq_emb": seq_emb}
