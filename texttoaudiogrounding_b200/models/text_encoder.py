"""Word-embedding text encoder — host-side mirror of reference models/text_encoder.py:14-43
(EmbeddingLayer) and :61-88 (EmbeddingAgg, aggregation="mean").  The gather + masked mean and
its scatter-add backward run in csrc/head.cu."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn as nn

import math

from ..ops import call
from . import nn_ops
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


class _AttnPoolFunction(torch.autograd.Function):
    """AttentionPooling.forward (reference models/text_encoder.py:51-58) as one kernel per direction."""

    @staticmethod
    def forward(ctx, x, lens, w, bias):
        B, N, D = x.shape
        out = torch.empty(B, D, device=x.device, dtype=torch.float32)
        weight = torch.empty(B, N, device=x.device, dtype=torch.float32)
        call("tag_attn_pool_fwd", x, lens, w, bias, out, weight, B, N, D)
        ctx.save_for_backward(x, w, weight)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, w, weight = ctx.saved_tensors
        B, N, D = x.shape
        d_x = torch.empty_like(x)
        d_w = torch.zeros_like(w)
        d_b = torch.zeros(1, device=x.device, dtype=torch.float32)
        call("tag_attn_pool_bwd", d_out.contiguous(), x, w, weight, d_x, d_w, d_b, B, N, D)
        return d_x, None, d_w, d_b


class AttentionPooling(nn.Module):
    """Mirror of reference models/text_encoder.py:46-58: ``fc`` = Linear(emb_dim, 1) holds the parameters (same
    state-dict keys ``attn.fc.{weight,bias}``); score, mask, softmax and the weighted sum run in csrc/head.cu."""

    def __init__(self, emb_dim):
        super().__init__()
        self.fc = nn.Linear(emb_dim, 1)

    def forward(self, x, lens):
        if not x.is_cuda:
            raise RuntimeError("AttentionPooling (B200) needs CUDA tensors: there is no CPU fallback")
        host_lens = torch.as_tensor(lens)
        if not host_lens.is_cuda and int(host_lens.max()) != x.shape[1]:
            # generate_length_mask(lens) is max(lens) wide; masked_fill then fails to broadcast (text_encoder.py:54-55)
            raise RuntimeError(f"The size of tensor a ({x.shape[1]}) must match the size of tensor b "
                               f"({int(host_lens.max())}) at non-singleton dimension 1")
        return _AttnPoolFunction.apply(x.float().contiguous(), lens_to_device(lens, x.device).contiguous(),
                                       self.fc.weight.reshape(-1), self.fc.bias)


class EmbeddingAgg(nn.Module):
    def __init__(self, vocab_size, embed_dim, pretrained_embedding: str = None,
                 freeze_embedding: bool = False, aggregation: str = "mean"):
        super().__init__()
        self.embedding = EmbeddingLayer(vocab_size, embed_dim, pretrained_embedding, freeze_embedding)
        self.embed_dim = self.embedding.embed_dim
        self.agg = aggregation
        if aggregation == "attention":
            self.attn = AttentionPooling(embed_dim)

    def forward(self, input_dict):
        if self.agg not in ("mean", "attention"):
            raise Exception(f"{self.agg} not supported")
        weight = self.embedding.core.weight
        if not weight.is_cuda:
            raise RuntimeError("EmbeddingAgg (B200) needs CUDA tensors: there is no CPU fallback")
        tokens = input_dict["text"].long().to(weight.device).contiguous()
        lens = lens_to_device(input_dict["text_len"], weight.device).contiguous()
        token_emb, seq_emb = _EmbedMeanFunction.apply(weight, tokens, lens)
        if self.agg == "attention":
            seq_emb = self.attn(token_emb, input_dict["text_len"])
        return {"token_emb": token_emb, "seq_emb": seq_emb}


class PositionalEncoding(nn.Module):
    """Sinusoidal table of reference models/text_encoder.py:128-146 (buffer ``pe`` [1, max_len, d_model]); the
    addition and its dropout are fused into the text-assemble kernel of SelfAttention."""

    def __init__(self, d_model, dropout, max_len=100):
        super().__init__()
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * -(math.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0))


class _TextAssembleFunction(torch.autograd.Function):
    """x[b] = dropout([cls; emb[text[b]]] + pe[:N+1])"""

    @staticmethod
    def forward(ctx, emb, cls, pe, tokens, dropout_p, seed):
        B, N = tokens.shape
        V, E = emb.shape
        out = torch.empty(B, N + 1, E, device=emb.device, dtype=torch.float32)
        call("tag_text_assemble_fwd", tokens, emb, cls, pe, out, B, N, E, V, float(dropout_p), seed, nn_ops.SEED_DEV)
        ctx.save_for_backward(tokens)
        ctx.cfg = (V, E, float(dropout_p), seed)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (tokens,) = ctx.saved_tensors
        V, E, p, seed = ctx.cfg
        B, N = tokens.shape
        d_emb = torch.zeros(V, E, device=d_out.device, dtype=torch.float32)
        d_cls = torch.zeros(1, 1, E, device=d_out.device, dtype=torch.float32)
        call("tag_text_assemble_bwd", tokens, d_out.contiguous(), d_emb, d_cls, B, N, E, V, p, seed, nn_ops.SEED_DEV)
        return d_emb, d_cls, None, None, None, None


class SelfAttention(nn.Module):
    """Transformer text encoder — mirror of reference models/text_encoder.py:240-268: word embedding with a [CLS]
    token and sinusoidal positions, one nn.MultiheadAttention layer with the padding mask; ``seq_emb`` is the [CLS]
    output, ``token_emb`` the rest.  Same constructor, parameters and state-dict keys (the nn.MultiheadAttention
    module holds the weights; its forward is replaced by csrc/attn.cu + the fp32 GEMMs)."""

    def __init__(self, vocab_size, embed_dim, num_heads, dropout=0.2, pretrained_embedding=None,
                 freeze_embedding=False) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.embedding = EmbeddingLayer(vocab_size, embed_dim, pretrained_embedding, freeze_embedding)
        self.pe = PositionalEncoding(embed_dim, dropout)
        self.mha = nn.MultiheadAttention(embed_dim, num_heads, dropout, batch_first=True)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))

    def forward(self, input_dict):
        weight = self.embedding.core.weight
        if not weight.is_cuda:
            raise RuntimeError("SelfAttention (B200) needs CUDA tensors: there is no CPU fallback")
        tokens = input_dict["text"].long().to(weight.device).contiguous()
        if tokens.shape[1] + 1 > self.pe.pe.shape[1]:
            raise RuntimeError(f"text longer than the positional table ({self.pe.pe.shape[1] - 1} tokens)")
        p = self.pe.dropout.p if self.training else 0.0
        x = _TextAssembleFunction.apply(weight, self.cls_token, self.pe.pe, tokens, p, nn_ops.next_seed())
        lens = (lens_to_device(input_dict["text_len"], weight.device) + 1).contiguous()
        x = nn_ops.multi_head_attention(self.mha, x, x, x, lens, self.training)
        return {"token_emb": x[:, 1:], "seq_emb": x[:, 0]}


class LaionClapEncoder(nn.Module):
    """CLAP text tower — mirror of reference models/text_encoder.py:311-327 (== models/hf_modeling_grounding.py:183-199):
    ``transformers`` ClapModel.text_model (RoBERTa-base layout) + ClapModel.text_projection; ``token_emb`` is the
    projection of every hidden state, ``seq_emb`` the L2-normalised projection of the pooler output.  The Hugging Face
    modules are kept as parameter containers (identical state-dict keys: ``model.*``, ``projection.*``); the forward
    pass runs on this repository's kernels: bf16 tcgen05 GEMMs for the 74 dense layers, csrc/attn.cu for the
    attention core, csrc/transformer.cu for embeddings / LayerNorm / GELU / normalisation.  Inference path (no
    autograd through the tower — the reference's released checkpoints keep it frozen).

    ``model_type`` is a Hugging Face model id / local path as in the reference, or a ``ClapConfig`` /
    ``ClapTextConfig`` for a randomly initialised tower (offline use)."""

    def __init__(self, model_type):
        super().__init__()
        from transformers import ClapConfig, ClapTextConfig
        if isinstance(model_type, (ClapConfig, ClapTextConfig)):
            from transformers.models.clap import modeling_clap as mc
            tcfg = model_type.text_config if isinstance(model_type, ClapConfig) else model_type
            self.tokenizer = None
            self.model = mc.ClapTextModel(tcfg)
            self.projection = mc.ClapProjectionLayer(tcfg)
            self.embed_dim = tcfg.projection_dim
        else:
            from transformers import ClapModel, ClapProcessor
            self.tokenizer = ClapProcessor.from_pretrained(model_type)
            model = ClapModel.from_pretrained(model_type)
            self.model = model.text_model
            self.projection = model.text_projection
            self.embed_dim = model.text_projection.config.projection_dim
        self._bf16 = {}
        # the tower is ~190 short launches per call (launch bound): from the third call with the same (batch, length)
        # the whole forward replays as one CUDA graph on static buffers
        self.use_graph = True
        self._graphs = {}
        self._seen = {}

    def refresh_operands(self) -> None:
        """Drop the cached bf16 copies of the dense weights and the captured graphs (call after loading / changing
        the weights)."""
        self._bf16 = {}
        self._graphs = {}
        self._seen = {}

    def _load_from_state_dict(self, *args, **kwargs):
        self.refresh_operands()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *a, **k):
        self.refresh_operands()
        return super()._apply(fn, *a, **k)

    def _dense(self, x: torch.Tensor, lin: nn.Linear, relu: bool = False) -> torch.Tensor:
        """x [R, K] fp32 -> x W^T + b [R, N] fp32 through a bf16 tensor-core GEMM (weights cast once)."""
        from .. import ops
        key = id(lin)
        wb = self._bf16.get(key)
        if wb is None:
            wb = ops.to_bf16(lin.weight.detach().float().contiguous())
            self._bf16[key] = wb
        R, K = x.shape
        N = lin.weight.shape[0]
        y = torch.empty(R, N, device=x.device, dtype=torch.float32)
        ops.annotate(f"fwd M={R} N={N} K={K}", 2.0 * R * N * K)
        call("tag_conv_tc_fwd", ops.to_bf16(x), wb, y, ops.F32, lin.bias.detach().float().contiguous(), int(relu), None,
             1, R, 1, K, N, 1)
        return y

    @torch.no_grad()
    def forward(self, input_dict):
        dev = self.model.embeddings.word_embeddings.weight.device
        if dev.type != "cuda":
            raise RuntimeError("LaionClapEncoder (B200) needs CUDA tensors: there is no CPU fallback")
        ids = input_dict["input_ids"].long().to(dev).contiguous()
        mask = input_dict["attention_mask"].long().to(dev).contiguous()
        key = tuple(ids.shape)
        if not self.use_graph or torch.cuda.is_current_stream_capturing():
            return self._forward_impl(ids, mask)
        entry = self._graphs.get(key)
        if entry is None:
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] < 3:
                return self._forward_impl(ids, mask)          # eager: also primes the bf16 weight copies
            s_ids, s_mask = ids.clone(), mask.clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._forward_impl(s_ids, s_mask)
            entry = self._graphs[key] = (graph, s_ids, s_mask, out)
        graph, s_ids, s_mask, out = entry
        s_ids.copy_(ids)
        s_mask.copy_(mask)
        graph.replay()
        return {k: v.clone() for k, v in out.items()}

    def _forward_impl(self, ids, mask):
        emb = self.model.embeddings
        dev = ids.device
        B, L = ids.shape
        cfg = self.model.config
        E, heads = cfg.hidden_size, cfg.num_attention_heads
        if L > 128:
            raise NotImplementedError("LaionClapEncoder (B200): at most 128 tokens per phrase")
        key_len = mask.sum(dim=-1).contiguous()          # right-padded batches (tokenizer(padding=True))
        x = torch.empty(B * L, E, device=dev, dtype=torch.float32)
        call("tag_roberta_embed_ln", ids, emb.word_embeddings.weight, emb.position_embeddings.weight,
             emb.token_type_embeddings.weight, emb.LayerNorm.weight, emb.LayerNorm.bias, x, B, L, E,
             emb.word_embeddings.weight.shape[0], emb.position_embeddings.weight.shape[0], emb.padding_idx,
             float(emb.LayerNorm.eps))
        for layer in self.model.encoder.layer:
            att = layer.attention
            q = self._dense(x, att.self.query).view(B, L, E)
            k = self._dense(x, att.self.key).view(B, L, E)
            v = self._dense(x, att.self.value).view(B, L, E)
            ctx = torch.empty_like(q)
            probs = torch.empty(B, heads, L, L, device=dev, dtype=torch.float32)
            call("tag_mha_core_fwd", q, k, v, key_len, ctx, probs, B, L, L, E, heads, 0.0, 0, None)
            a = self._dense(ctx.view(B * L, E), att.output.dense)
            x1 = torch.empty_like(x)
            call("tag_add_layernorm", a, x, att.output.LayerNorm.weight, att.output.LayerNorm.bias, x1, B * L, E,
                 float(att.output.LayerNorm.eps))
            h = self._dense(x1, layer.intermediate.dense)
            call("tag_unary_f32", h, h, h.numel(), 0)                       # exact GELU
            o = self._dense(h, layer.output.dense)
            x = torch.empty_like(x1)
            call("tag_add_layernorm", o, x1, layer.output.LayerNorm.weight, layer.output.LayerNorm.bias, x, B * L, E,
                 float(layer.output.LayerNorm.eps))
        first = x.view(B, L, E)[:, 0].contiguous()
        pooled = self._dense(first, self.model.pooler.dense)
        call("tag_unary_f32", pooled, pooled, pooled.numel(), 1)           # tanh
        token_emb = self._dense(self._dense(x, self.projection.linear1, relu=True), self.projection.linear2)
        seq = self._dense(self._dense(pooled, self.projection.linear1, relu=True), self.projection.linear2)
        seq_emb = torch.empty_like(seq)
        call("tag_l2_normalize", seq, seq_emb, B, seq.shape[1], 1e-12)
        return {"seq_emb": seq_emb, "token_emb": token_emb.view(B, L, -1)}
