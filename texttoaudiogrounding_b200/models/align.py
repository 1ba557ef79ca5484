"""All-pairs audio<->text alignment — host-side mirror of reference models/align.py:7-31 (DotProduct).

The reference returns the 4-D matrix sim[i, j, t, n] for every (clip i, text j) pair and hands it to a
``sim_pooling`` module.  Here ``DotProduct.forward`` returns a :class:`PairwiseSim` — the operands of that matrix,
not its Ba*Bt*T*N values — and the sim_pooling modules (models/sim_pooling.py) run ONE fused kernel
(csrc/align.cu) that forms the dot products and pools over the frames in registers.  ``PairwiseSim.materialize()``
gives the reference's tensor when a caller asks for it (``output_matrix``)."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from .. import ops
from ..ops import call
from .utils import lens_to_device

AUDIO_POOL = {"mean": 0, "max": 1, "linear_softmax": 2, "exp_softmax": 3}
TEXT_POOL = {"mean": 0, "sum": 1, "max": 2, "meansum": 3}


TC_MIN_ROWS = 1024      # Ba*T from which the split-bf16 tensor-core route is used


def _pad_rows(text2d: torch.Tensor, mult: int = 64) -> torch.Tensor:
    C, D = text2d.shape
    Cpad = (C + mult - 1) // mult * mult
    if Cpad == C:
        return text2d.contiguous()
    out = torch.zeros(Cpad, D, device=text2d.device, dtype=torch.float32)
    out[:C].copy_(text2d)
    return out


def _use_tc(rows: int) -> bool:
    return ops.USE_TC and rows >= TC_MIN_ROWS


def _bf16(*shape, device):
    return torch.empty(*shape, device=device, dtype=torch.bfloat16)


def _tc_logits(audio, textp):
    """[Ba*T, Cpad] fp32 dot products through ONE split-bf16 tcgen05 GEMM of depth 3D (csrc/split.cu)."""
    Ba, T, D = audio.shape
    M, Cpad = Ba * T, textp.shape[0]
    a_k, x_k = _bf16(M, 3 * D, device=audio.device), _bf16(Cpad, 3 * D, device=audio.device)
    call("tag_split_bf16x3", audio, a_k, M, D, 0, 0)
    call("tag_split_bf16x3", textp, x_k, Cpad, D, 1, 0)
    logits = torch.empty(M, Cpad, device=audio.device, dtype=torch.float32)
    ops.annotate(f"fwd M={M} N={Cpad} K={3 * D}", 2.0 * M * Cpad * 3 * D)
    call("tag_conv_tc_fwd", a_k, x_k, logits, ops.F32, None, 0, None, 1, M, 1, 3 * D, Cpad, 1)
    return logits


class _AlignPoolFunction(torch.autograd.Function):
    """out[i, j] = text-pool_n( frame-pool_t( sim[i, j, t, n] ) ) without materialising sim.  Two routes: small
    batches run the fused CUDA-core kernel (scores + pooling in registers, backward recomputes); from TC_MIN_ROWS rows
    the scores come from a split-bf16 tcgen05 GEMM (fp32-level accuracy) and the logits [Ba*T, Cpad] are kept for
    backward, whose two gradient GEMMs run on the tensor cores as well."""

    @staticmethod
    def forward(ctx, audio, text, audio_len, text_len, a_mode, t_mode, scale):
        Ba, T, D = audio.shape
        Bt, N, _ = text.shape
        tc = _use_tc(Ba * T)
        textp = _pad_rows(text.reshape(Bt * N, D), 128 if tc else 64)
        Cpad = textp.shape[0]
        colpool = torch.empty(Ba, Cpad, device=audio.device, dtype=torch.float32)
        aux = torch.empty_like(colpool)
        logits = None
        if tc:
            logits = _tc_logits(audio, textp)
            call("tag_align_logits_fwd", logits, audio_len, a_mode, None, colpool, aux, Ba, T, Bt, N, Cpad, scale)
        else:
            ops.annotate(f"align fwd Ba={Ba} T={T} C={Bt * N}", 2.0 * Ba * T * Cpad * D)
            call("tag_align_pool_fwd", audio, textp, audio_len, a_mode, None, colpool, aux, Ba, T, Bt, N, Cpad, D,
                 scale)
        out = torch.empty(Ba, Bt, device=audio.device, dtype=torch.float32)
        call("tag_align_text_pool_fwd", colpool, text_len, t_mode, out, Ba, Bt, N, Cpad)
        ctx.save_for_backward(audio, textp, audio_len, text_len, colpool, aux, logits)
        ctx.cfg = (a_mode, t_mode, scale, Bt, N)
        return out

    @staticmethod
    def backward(ctx, d_out):
        audio, textp, audio_len, text_len, colpool, aux, logits = ctx.saved_tensors
        a_mode, t_mode, scale, Bt, N = ctx.cfg
        Ba, T, D = audio.shape
        M, Cpad = Ba * T, textp.shape[0]
        dev = audio.device
        d_colpool = torch.zeros(Ba, Cpad, device=dev, dtype=torch.float32)
        call("tag_align_text_pool_bwd", d_out.contiguous(), colpool, text_len, t_mode, d_colpool, Ba, Bt, N, Cpad)
        d_audio = d_text = None
        if logits is not None:
            g_k, g_p = _bf16(M, 3 * Cpad, device=dev), _bf16(2, M, Cpad, device=dev)
            call("tag_align_logits_bwd", logits, audio_len, a_mode, d_colpool, colpool, aux, g_k, g_p, Ba, T, Cpad,
                 scale)
            if ctx.needs_input_grad[0]:
                xt_k = _bf16(D, 3 * Cpad, device=dev)
                call("tag_split_bf16x3", textp, xt_k, D, Cpad, 1, 1)          # text^T as the [D][3*Cpad] B operand
                d_audio = torch.empty(Ba, T, D, device=dev, dtype=torch.float32)
                ops.annotate(f"fwd M={M} N={D} K={3 * Cpad}", 2.0 * M * D * 3 * Cpad)
                call("tag_conv_tc_fwd", g_k, xt_k, d_audio, ops.F32, None, 0, None, 1, M, 1, 3 * Cpad, D, 1)
            if ctx.needs_input_grad[1]:
                a_p = _bf16(2, M, D, device=dev)
                call("tag_split_bf16x3", audio, a_p, M, D, 2, 0)
                d_textp = torch.zeros(Cpad, D, device=dev, dtype=torch.float32)
                splits = ops.tc_wgrad_splits(1, M, 1, D, Cpad, 1)
                for gh, ah in ((0, 0), (0, 1), (1, 0)):                        # hi.hi + hi.lo + lo.hi
                    ops.annotate(f"wgrad P={M} Cout={Cpad} K={D}", 2.0 * M * Cpad * D)
                    call("tag_conv_tc_wgrad", g_p[gh], a_p[ah], d_textp, 1, M, 1, D, Cpad, 1, splits)
                d_text = d_textp[:Bt * N].view(Bt, N, D)
            return d_audio, d_text, None, None, None, None, None
        G = torch.empty(M, Cpad, device=dev, dtype=torch.float32)
        ops.annotate(f"align bwd Ba={Ba} T={T} C={Bt * N}", 2.0 * Ba * T * Cpad * D)
        call("tag_align_pool_bwd", audio, textp, audio_len, a_mode, d_colpool, colpool, aux, G, Ba, T, Cpad, D, scale)
        if ctx.needs_input_grad[0]:
            # d_audio [Ba*T, D] = G [Ba*T, Cpad] x text [Cpad, D]: the 1x1 "conv" GEMM wants the weight as [D][Cpad]
            text_t = torch.empty(D * Cpad, device=dev, dtype=torch.float32)
            call("tag_weight_flip_transpose", textp, text_t, Cpad, D, 1)
            d_audio = torch.empty(Ba, T, D, device=dev, dtype=torch.float32)
            ops.annotate(f"fwd M={Ba * T} N={D} K={Cpad}", 2.0 * Ba * T * Cpad * D)
            call("tag_conv_fwd", G, ops.F32, text_t, d_audio, ops.F32, None, 0, None, 1, Ba * T, 1, Cpad, D, 1)
        if ctx.needs_input_grad[1]:
            # d_text [Cpad, D] = G^T x audio (split over the Ba*T rows, accumulated with atomics)
            d_textp = torch.zeros(Cpad, D, device=dev, dtype=torch.float32)
            ops.annotate(f"wgrad P={Ba * T} Cout={Cpad} K={D}", 2.0 * Ba * T * Cpad * D)
            call("tag_conv_wgrad", G, ops.F32, audio, ops.F32, d_textp, 1, Ba * T, 1, D, Cpad, 1,
                 ops.wgrad_splits(Ba * T, D, Cpad, 1))
            d_text = d_textp[:Bt * N].view(Bt, N, D)
        return d_audio, d_text, None, None, None, None, None


class PairwiseSim:
    """Deferred sim[i, j, t, n] = clamp(sigmoid(scale * <audio[i,t,:], text[j,n,:]>), 1e-7, 1)."""

    def __init__(self, audio: torch.Tensor, text: torch.Tensor, scale: float):
        if not audio.is_cuda:
            raise RuntimeError("align.DotProduct (B200) needs CUDA tensors: there is no CPU fallback")
        a_bs, n_seg, a_dim = audio.size()
        t_bs, n_txt, t_dim = text.size()
        assert a_bs == t_bs
        assert a_dim == t_dim
        self.audio = audio.float().contiguous()
        self.text = text.float().contiguous()
        self.scale = scale

    def size(self, dim=None):
        shape = torch.Size((self.audio.size(0), self.text.size(0), self.audio.size(1), self.text.size(1)))
        return shape if dim is None else shape[dim]

    @property
    def shape(self):
        return self.size()

    def pool(self, audio_len, text_len, audio_pool: str, text_pool: str) -> torch.Tensor:
        dev = self.audio.device
        return _AlignPoolFunction.apply(self.audio, self.text, lens_to_device(audio_len, dev).contiguous(),
                                        lens_to_device(text_len, dev).contiguous(), AUDIO_POOL[audio_pool],
                                        TEXT_POOL[text_pool], self.scale)

    @torch.no_grad()
    def materialize(self) -> torch.Tensor:
        """The reference's [Ba, Bt, T, N] tensor (not differentiable: it is an inspection output)."""
        Ba, T, D = self.audio.shape
        Bt, N, _ = self.text.shape
        tc = _use_tc(Ba * T)
        textp = _pad_rows(self.text.reshape(Bt * N, D), 128 if tc else 64)
        Cpad = textp.shape[0]
        dev = self.audio.device
        sim = torch.empty(Ba, Bt, T, N, device=dev, dtype=torch.float32)
        colpool = torch.empty(Ba, Cpad, device=dev, dtype=torch.float32)
        aux = torch.empty_like(colpool)
        full = torch.full((Ba,), T, device=dev, dtype=torch.long)
        if tc:
            call("tag_align_logits_fwd", _tc_logits(self.audio, textp), full, 0, sim, colpool, aux, Ba, T, Bt, N, Cpad,
                 self.scale)
        else:
            call("tag_align_pool_fwd", self.audio, textp, full, 0, sim, colpool, aux, Ba, T, Bt, N, Cpad, D,
                 self.scale)
        return sim


class DotProduct(nn.Module):
    def __init__(self, l2norm=False, scaled=False) -> None:
        super().__init__()
        self.l2norm = l2norm
        self.scaled = scaled
        if l2norm:
            raise NotImplementedError("align.DotProduct(l2norm=True) is not on the B200 path (no eg_config uses it)")

    def forward(self, audio: torch.Tensor, text: torch.Tensor, **kwargs) -> PairwiseSim:
        scale = 1.0 / math.sqrt(audio.size(-1)) if self.scaled else 1.0
        return PairwiseSim(audio, text, scale)
