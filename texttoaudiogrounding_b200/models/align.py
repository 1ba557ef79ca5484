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


def _pad_rows(text2d: torch.Tensor) -> torch.Tensor:
    C, D = text2d.shape
    Cpad = (C + 63) // 64 * 64
    if Cpad == C:
        return text2d.contiguous()
    out = torch.zeros(Cpad, D, device=text2d.device, dtype=torch.float32)
    out[:C].copy_(text2d)
    return out


class _AlignPoolFunction(torch.autograd.Function):
    """out[i, j] = text-pool_n( frame-pool_t( sim[i, j, t, n] ) ) without materialising sim."""

    @staticmethod
    def forward(ctx, audio, text, audio_len, text_len, a_mode, t_mode, scale):
        Ba, T, D = audio.shape
        Bt, N, _ = text.shape
        textp = _pad_rows(text.reshape(Bt * N, D))
        Cpad = textp.shape[0]
        colpool = torch.empty(Ba, Cpad, device=audio.device, dtype=torch.float32)
        aux = torch.empty_like(colpool)
        ops.annotate(f"align fwd Ba={Ba} T={T} C={Bt * N}", 2.0 * Ba * T * Cpad * D)
        call("tag_align_pool_fwd", audio, textp, audio_len, a_mode, None, colpool, aux, Ba, T, Bt, N, Cpad, D, scale)
        out = torch.empty(Ba, Bt, device=audio.device, dtype=torch.float32)
        call("tag_align_text_pool_fwd", colpool, text_len, t_mode, out, Ba, Bt, N, Cpad)
        ctx.save_for_backward(audio, textp, audio_len, text_len, colpool, aux)
        ctx.cfg = (a_mode, t_mode, scale, Bt, N)
        return out

    @staticmethod
    def backward(ctx, d_out):
        audio, textp, audio_len, text_len, colpool, aux = ctx.saved_tensors
        a_mode, t_mode, scale, Bt, N = ctx.cfg
        Ba, T, D = audio.shape
        Cpad = textp.shape[0]
        dev = audio.device
        d_colpool = torch.zeros(Ba, Cpad, device=dev, dtype=torch.float32)
        call("tag_align_text_pool_bwd", d_out.contiguous(), colpool, text_len, t_mode, d_colpool, Ba, Bt, N, Cpad)
        G = torch.empty(Ba * T, Cpad, device=dev, dtype=torch.float32)
        ops.annotate(f"align bwd Ba={Ba} T={T} C={Bt * N}", 2.0 * Ba * T * Cpad * D)
        call("tag_align_pool_bwd", audio, textp, audio_len, a_mode, d_colpool, colpool, aux, G, Ba, T, Cpad, D, scale)
        d_audio = d_text = None
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
        textp = _pad_rows(self.text.reshape(Bt * N, D))
        Cpad = textp.shape[0]
        dev = self.audio.device
        sim = torch.empty(Ba, Bt, T, N, device=dev, dtype=torch.float32)
        colpool = torch.empty(Ba, Cpad, device=dev, dtype=torch.float32)
        aux = torch.empty_like(colpool)
        full = torch.full((Ba,), T, device=dev, dtype=torch.long)
        call("tag_align_pool_fwd", self.audio, textp, full, 0, sim, colpool, aux, Ba, T, Bt, N, Cpad, D, self.scale)
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
