"""Autograd bindings of the small fp32 building blocks shared by the attention-type heads:
nn.Linear, the nn.MultiheadAttention core, additive attention, sigmoid gating.  Every forward/backward is a C-ABI
call (csrc/attn.cu, csrc/conv_simt.cu); nothing here computes with torch ops."""
from __future__ import annotations

import itertools

import torch

from .. import ops
from ..ops import call

_SEEDS = itertools.count(0x5EED)
# device-side addend of every dropout seed (a 1-element int64/uint64 CUDA tensor, or None).  A graph-replayed train step
# points it at its step counter so that replays draw fresh masks although the host seeds are frozen in the graph.
SEED_DEV = None


def next_seed() -> int:
    """A fresh dropout seed per call (counter-based RNG: the same seed regenerates the mask in backward)."""
    return next(_SEEDS)


def _need_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} (B200) needs CUDA tensors: there is no CPU fallback")


TC_MIN_ROWS = 1024      # rows from which F.linear runs as split-bf16 tcgen05 GEMMs (csrc/split.cu)


def _bf16(*shape, device):
    return torch.empty(*shape, device=device, dtype=torch.bfloat16)


class _LinearFunction(torch.autograd.Function):
    """y [R, Cout] = x [R, Cin] W^T + b — F.linear; Cin % 64 == 0, Cout % 64 == 0.  Small R: fp32 CUDA-core GEMMs
    (conv_simt.cu).  R >= TC_MIN_ROWS: every operand is split into bf16 hi + lo and the three products
    hi.hi + hi.lo + lo.hi run as ONE bf16 tcgen05 GEMM of depth 3K with fp32 accumulation (fp32-level accuracy)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        R, Cin = x.shape
        Cout = weight.shape[0]
        y = torch.empty(R, Cout, device=x.device, dtype=torch.float32)
        ctx.tc = ops.USE_TC and R >= TC_MIN_ROWS and Cin >= 128
        if ctx.tc:
            x_k, w_k = _bf16(R, 3 * Cin, device=x.device), _bf16(Cout, 3 * Cin, device=x.device)
            call("tag_split_bf16x3", x, x_k, R, Cin, 0, 0)
            call("tag_split_bf16x3", weight, w_k, Cout, Cin, 1, 0)
            ops.annotate(f"fwd M={R} N={Cout} K={3 * Cin}", 2.0 * R * Cout * 3 * Cin)
            call("tag_conv_tc_fwd", x_k, w_k, y, ops.F32, bias, 0, None, 1, R, 1, 3 * Cin, Cout, 1)
        else:
            ops.annotate(f"fwd M={R} N={Cout} K={Cin}", 2.0 * R * Cout * Cin)
            call("tag_conv_fwd", x, ops.F32, weight, y, ops.F32, bias, 0, None, 1, R, 1, Cin, Cout, 1)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        R, Cin = x.shape
        Cout = weight.shape[0]
        dev = x.device
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(R, Cin, device=dev, dtype=torch.float32)
            if ctx.tc:
                dy_k, wt_k = _bf16(R, 3 * Cout, device=dev), _bf16(Cin, 3 * Cout, device=dev)
                call("tag_split_bf16x3", dy, dy_k, R, Cout, 0, 0)
                call("tag_split_bf16x3", weight, wt_k, Cin, Cout, 1, 1)           # W^T as the [Cin][3*Cout] operand
                ops.annotate(f"fwd M={R} N={Cin} K={3 * Cout}", 2.0 * R * Cin * 3 * Cout)
                call("tag_conv_tc_fwd", dy_k, wt_k, dx, ops.F32, None, 0, None, 1, R, 1, 3 * Cout, Cin, 1)
            else:
                wt = torch.empty(Cin * Cout, device=dev, dtype=torch.float32)
                call("tag_weight_flip_transpose", weight, wt, Cout, Cin, 1)
                ops.annotate(f"fwd M={R} N={Cin} K={Cout}", 2.0 * R * Cout * Cin)
                call("tag_conv_fwd", dy, ops.F32, wt, dx, ops.F32, None, 0, None, 1, R, 1, Cout, Cin, 1)
        if ctx.needs_input_grad[1]:
            dw = torch.zeros(Cout, Cin, device=dev, dtype=torch.float32)
            if ctx.tc:
                dy_p, x_p = _bf16(2, R, Cout, device=dev), _bf16(2, R, Cin, device=dev)
                call("tag_split_bf16x3", dy, dy_p, R, Cout, 2, 0)
                call("tag_split_bf16x3", x, x_p, R, Cin, 2, 0)
                splits = ops.tc_wgrad_splits(1, R, 1, Cin, Cout, 1)
                for gh, xh in ((0, 0), (0, 1), (1, 0)):
                    ops.annotate(f"wgrad P={R} Cout={Cout} K={Cin}", 2.0 * R * Cout * Cin)
                    call("tag_conv_tc_wgrad", dy_p[gh], x_p[xh], dw, 1, R, 1, Cin, Cout, 1, splits)
            else:
                ops.annotate(f"wgrad P={R} Cout={Cout} K={Cin}", 2.0 * R * Cout * Cin)
                call("tag_conv_wgrad", dy, ops.F32, x, ops.F32, dw, 1, R, 1, Cin, Cout, 1,
                     ops.wgrad_splits(R, Cin, Cout, 1))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = torch.zeros(Cout, device=dev, dtype=torch.float32)
            call("tag_colsum", dy, ops.F32, R, Cout, db)
        return dx, dw, db


def linear(x: torch.Tensor, weight: torch.Tensor, bias=None) -> torch.Tensor:
    _need_cuda(x, "linear")
    if weight.shape[1] % 64 != 0 or weight.shape[0] % 64 != 0:
        raise NotImplementedError("linear (B200): in/out features must be multiples of 64")
    lead = x.shape[:-1]
    y = _LinearFunction.apply(x.reshape(-1, x.shape[-1]).float().contiguous(), weight.float().contiguous(),
                              None if bias is None else bias.float().contiguous())
    return y.view(*lead, weight.shape[0])


class _MhaCoreFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, key_len, heads, dropout_p, seed):
        B, Lq, E = q.shape
        Lk = k.shape[1]
        out = torch.empty_like(q)
        probs = torch.empty(B, heads, Lq, Lk, device=q.device, dtype=torch.float32)
        call("tag_mha_core_fwd", q, k, v, key_len, out, probs, B, Lq, Lk, E, heads, float(dropout_p), seed, SEED_DEV)
        ctx.save_for_backward(q, k, v, probs, key_len)
        ctx.cfg = (heads, float(dropout_p), seed)
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, probs, key_len = ctx.saved_tensors
        heads, p, seed = ctx.cfg
        B, Lq, E = q.shape
        Lk = k.shape[1]
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        call("tag_mha_core_bwd", d_out.contiguous(), q, k, v, probs, key_len, dq, dk, dv, B, Lq, Lk, E, heads, p, seed,
             SEED_DEV)
        return dq, dk, dv, None, None, None, None


def multi_head_attention(mha: torch.nn.MultiheadAttention, query, key, value, key_len, training: bool):
    """nn.MultiheadAttention.forward(query, key, value, key_padding_mask = arange(Lk) >= key_len[:, None]) for a
    batch_first module with a packed in_proj (kdim == vdim == embed_dim); returns the attention output only."""
    if not mha._qkv_same_embed_dim or not mha.batch_first or mha.bias_k is not None or mha.add_zero_attn:
        raise NotImplementedError("multi_head_attention (B200): packed in_proj, batch_first, no bias_k / zero_attn")
    _need_cuda(query, "MultiheadAttention")
    E = mha.embed_dim
    W, b = mha.in_proj_weight, mha.in_proj_bias
    bq, bk, bv = (None, None, None) if b is None else (b[:E], b[E:2 * E], b[2 * E:])
    q = linear(query, W[:E], bq)
    k = linear(key, W[E:2 * E], bk)
    v = linear(value, W[2 * E:], bv)
    p = mha.dropout if training else 0.0
    ctx = _MhaCoreFunction.apply(q, k, v, key_len, mha.num_heads, p, next_seed())
    return linear(ctx, mha.out_proj.weight, mha.out_proj.bias)


class _AdditiveAttnFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, hq, hk, v, kv, q_len, kv_len):
        B, T, E = hq.shape
        N = hk.shape[1]
        attn = torch.empty(B, T, N, device=hq.device, dtype=torch.float32)
        out = torch.empty(B, T, E, device=hq.device, dtype=torch.float32)
        call("tag_additive_attn_fwd", hq, hk, v, kv, q_len, kv_len, attn, out, B, T, N, E)
        ctx.save_for_backward(hq, hk, v, kv, attn, q_len, kv_len)
        return out

    @staticmethod
    def backward(ctx, d_out):
        hq, hk, v, kv, attn, q_len, kv_len = ctx.saved_tensors
        B, T, E = hq.shape
        N = hk.shape[1]
        d_hq = torch.empty_like(hq)
        d_hk, d_v, d_kv = torch.zeros_like(hk), torch.zeros_like(v), torch.zeros_like(kv)
        call("tag_additive_attn_bwd", d_out.contiguous(), hq, hk, v, kv, attn, q_len, kv_len, d_hq, d_hk, d_v, d_kv,
             B, T, N, E)
        return d_hq, d_hk, d_v, d_kv, None, None


def additive_attention(hq, hk, v, kv, q_len, kv_len):
    return _AdditiveAttnFunction.apply(hq.contiguous(), hk.contiguous(), v.float().contiguous(), kv.float().contiguous(),
                                       q_len, kv_len)


class _SigmoidGateFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, z):
        out = torch.empty_like(x)
        call("tag_sigmoid_gate_fwd", x, z, out, x.numel())
        ctx.save_for_backward(x, z)
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, z = ctx.saved_tensors
        dx, dz = torch.empty_like(x), torch.empty_like(z)
        call("tag_sigmoid_gate_bwd", d_out.contiguous(), x, z, dx, dz, x.numel())
        return dx, dz


def sigmoid_gate(x, z):
    """x * sigmoid(z)"""
    return _SigmoidGateFunction.apply(x.float().contiguous(), z.float().contiguous())


class _RowDotSigmoidFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, x, scale):
        R, E = a.shape
        sim = torch.empty(R, device=a.device, dtype=torch.float32)
        call("tag_rowdot_sigmoid_fwd", a, x, sim, R, E, scale)
        ctx.save_for_backward(a, x, sim)
        ctx.scale = scale
        return sim

    @staticmethod
    def backward(ctx, d_sim):
        a, x, sim = ctx.saved_tensors
        R, E = a.shape
        da, dx = torch.empty_like(a), torch.empty_like(x)
        call("tag_rowdot_sigmoid_bwd", d_sim.contiguous(), sim, a, x, da, dx, R, E, ctx.scale)
        return da, dx, None


def rowdot_sigmoid(a, x, scale):
    """clamp(sigmoid(scale * <a[..., :], x[..., :]>), 1e-7, 1) over the last dimension"""
    lead = a.shape[:-1]
    E = a.shape[-1]
    return _RowDotSigmoidFunction.apply(a.reshape(-1, E).float().contiguous(), x.reshape(-1, E).float().contiguous(),
                                        scale).view(*lead)


class _LnLinearSigmoidFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, audio, attn_out, gamma, beta, w, bias, eps, dropout_p, seed):
        R, E = audio.shape
        prob = torch.empty(R, device=audio.device, dtype=torch.float32)
        stat = torch.empty(R, 2, device=audio.device, dtype=torch.float32)
        call("tag_ln_linear_sigmoid_fwd", audio, attn_out, gamma, beta, w, bias, prob, stat, R, E, float(eps),
             float(dropout_p), seed, SEED_DEV)
        ctx.save_for_backward(audio, attn_out, gamma, beta, w, prob, stat)
        ctx.cfg = (float(dropout_p), seed)
        return prob

    @staticmethod
    def backward(ctx, d_prob):
        audio, attn_out, gamma, beta, w, prob, stat = ctx.saved_tensors
        p, seed = ctx.cfg
        R, E = audio.shape
        d_audio, d_attn = torch.empty_like(audio), torch.empty_like(attn_out)
        d_gamma, d_beta, d_w = torch.zeros_like(gamma), torch.zeros_like(beta), torch.zeros_like(w)
        d_bias = torch.zeros(1, device=audio.device, dtype=torch.float32)
        call("tag_ln_linear_sigmoid_bwd", d_prob.contiguous(), prob, audio, attn_out, gamma, beta, w, stat, d_audio,
             d_attn, d_gamma, d_beta, d_w, d_bias, R, E, p, seed, SEED_DEV)
        return d_audio, d_attn, d_gamma, d_beta, d_w, d_bias, None, None, None


def ln_linear_sigmoid(audio, attn_out, norm: torch.nn.LayerNorm, lin: torch.nn.Linear, dropout_p: float):
    """sigmoid(lin(norm(audio + dropout(attn_out)))).squeeze(-1) for lin = Linear(E, 1)"""
    lead = audio.shape[:-1]
    E = audio.shape[-1]
    out = _LnLinearSigmoidFunction.apply(
        audio.reshape(-1, E).float().contiguous(), attn_out.reshape(-1, E).float().contiguous(),
        norm.weight.float().contiguous(), norm.bias.float().contiguous(), lin.weight.reshape(-1).float().contiguous(),
        lin.bias.float().contiguous(), norm.eps, dropout_p, next_seed())
    return out.view(*lead)


class _UpsampleLinearFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, size):
        outer, T, inner = x.shape
        y = torch.empty(outer, size, inner, device=x.device, dtype=torch.float32)
        call("tag_upsample_linear_fwd", x, y, outer, T, size, inner)
        ctx.cfg = (outer, T, inner, size)
        return y

    @staticmethod
    def backward(ctx, dy):
        outer, T, inner, size = ctx.cfg
        dx = torch.empty(outer, T, inner, device=dy.device, dtype=torch.float32)
        call("tag_upsample_linear_bwd", dy.contiguous(), dx, outer, T, size, inner)
        return dx, None


def upsample_linear(x: torch.Tensor, size: int) -> torch.Tensor:
    """F.interpolate(x, size=size, mode="linear", align_corners=False) along dim 1 of x [B, T] or [B, T, n]
    (the reference interpolates [B, 1, T] / [B, n, T] views, models/audio_text_model.py:90-97, 216-223)."""
    _need_cuda(x, "upsample")
    squeeze = x.dim() == 2
    x3 = x.unsqueeze(-1) if squeeze else x
    y = _UpsampleLinearFunction.apply(x3.float().contiguous(), int(size))
    return y.squeeze(-1) if squeeze else y
