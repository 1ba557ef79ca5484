"""FrameBceLoss / ClipBceLoss / ClipFrameBceLoss — host-side mirrors of reference losses.py:11-43 and :186-210.
Masked mean BCE over the frames t < length[b] (or over all clip-level probabilities); forward and
d(loss)/d(prob) in one kernel (csrc/head.cu)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn

from .ops import call


class _FrameBceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, frame_sim, label, length):
        B, T = frame_sim.shape
        loss = torch.empty((), device=frame_sim.device, dtype=torch.float32)
        d_sim = torch.empty(B, T, device=frame_sim.device, dtype=torch.float32)
        call("tag_frame_bce", frame_sim, frame_sim.stride(0), label, label.stride(0), length, B, T,
             loss, d_sim, T, 1.0)
        ctx.save_for_backward(d_sim)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        (d_sim,) = ctx.saved_tensors
        return d_sim * d_loss, None, None


def frame_bce(frame_sim: torch.Tensor, label: torch.Tensor, length) -> torch.Tensor:
    if frame_sim.ndim == 3 and frame_sim.size(2) == 1:
        frame_sim = frame_sim.squeeze(2)
    if frame_sim.ndim == 3:
        # [B, T, n] with the mask expanded over n (losses.py:26-35): element (b, t, j) is valid iff t < length[b],
        # i.e. iff its flat index t*n + j < length[b]*n
        n = frame_sim.size(2)
        B = frame_sim.size(0)
        length3 = torch.as_tensor(length).to(torch.long) * n
        label = label.to(device=frame_sim.device, dtype=torch.float32)
        return frame_bce(frame_sim.contiguous().view(B, -1), label.contiguous().view(B, -1), length3)
    if not frame_sim.is_cuda:
        raise RuntimeError("FrameBceLoss (B200) needs CUDA tensors: there is no CPU fallback")
    length_host = torch.as_tensor(length)
    T = frame_sim.size(1)
    if not length_host.is_cuda:
        # same failure as the reference's `loss *= mask` when the mask is narrower than the loss
        if int(length_host.max()) != T and int(length_host.max()) != 1:
            raise RuntimeError(f"The size of tensor a ({T}) must match the size of tensor b "
                               f"({int(length_host.max())}) at non-singleton dimension 1")
    if frame_sim.stride(1) != 1:
        frame_sim = frame_sim.contiguous()
    label = label.to(device=frame_sim.device, dtype=torch.float32)
    if label.stride(1) != 1:
        label = label.contiguous()
    length_dev = length_host.to(device=frame_sim.device, dtype=torch.long).contiguous()
    return _FrameBceFunction.apply(frame_sim.float(), label, length_dev)


class FrameBceLoss(nn.Module):
    def forward(self, output: Dict):
        return frame_bce(output["frame_sim"], output["label"], output["length"])

    def forward_tensor(self, frame_sim, label, length):
        return frame_bce(frame_sim, label, length)


def clip_bce(prob: torch.Tensor, label: torch.Tensor) -> torch.Tensor:
    """F.binary_cross_entropy(prob, label) (mean over all elements) of clip-level probabilities [B, n]."""
    if not prob.is_cuda:
        raise RuntimeError("ClipBceLoss (B200) needs CUDA tensors: there is no CPU fallback")
    p2 = prob.float().contiguous().view(prob.shape[0], -1)
    l2 = label.to(device=prob.device, dtype=torch.float32).contiguous().view(prob.shape[0], -1)
    full = torch.full((p2.shape[0],), p2.shape[1], device=prob.device, dtype=torch.long)
    return _FrameBceFunction.apply(p2, l2, full)


class ClipBceLoss(nn.Module):
    def forward(self, output: Dict):
        return clip_bce(output["clip_sim"], output["label"])

    def forward_tensor(self, prob, label):
        return clip_bce(prob, label)


class ClipFrameBceLoss(nn.Module):
    def __init__(self, frame_weight, clip_label_key="weak_label", clip_prob_key="clip_sim",
                 frame_label_key="strong_label", frame_prob_key="frame_sim"):
        super().__init__()
        self.clip_loss_fn = ClipBceLoss()
        self.frame_loss_fn = FrameBceLoss()
        self.frame_weight = frame_weight
        self.clip_label_key = clip_label_key
        self.clip_prob_key = clip_prob_key
        self.frame_label_key = frame_label_key
        self.frame_prob_key = frame_prob_key

    def forward(self, output: Dict):
        return (1 - self.frame_weight) * self.clip_loss_fn.forward_tensor(
            output[self.clip_prob_key], output[self.clip_label_key]) + \
            self.frame_weight * self.frame_loss_fn.forward_tensor(
                output[self.frame_prob_key], output[self.frame_label_key], output["length"])


class _MaxMarginRankFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim, margin, lamda1, fix_norm):
        n = sim.shape[0]
        loss = torch.empty((), device=sim.device, dtype=torch.float32)
        d_sim = torch.empty(n, n, device=sim.device, dtype=torch.float32)
        call("tag_max_margin_rank", sim, n, float(margin), float(lamda1), int(bool(fix_norm)), loss, d_sim)
        ctx.save_for_backward(d_sim)
        return loss

    @staticmethod
    def backward(ctx, d_loss):
        (d_sim,) = ctx.saved_tensors
        return d_sim * d_loss, None, None, None


class MaxMarginRankingLoss(nn.Module):
    """Bidirectional max-margin ranking over the clip x caption similarity matrix — mirror of reference
    losses.py:226-264 (same constructor; loss and its gradient from one kernel, csrc/align.cu)."""

    def __init__(self, margin=1, fix_norm=True, lamda1=1, sim_key="sim"):
        super().__init__()
        self.fix_norm = fix_norm
        self.margin = margin
        self.lamda1 = lamda1
        self.sim_key = sim_key

    def forward(self, x):
        x = x[self.sim_key]
        if not x.is_cuda:
            raise RuntimeError("MaxMarginRankingLoss (B200) needs CUDA tensors: there is no CPU fallback")
        if x.ndim != 2 or x.shape[0] != x.shape[1]:
            raise RuntimeError(f"expected a square similarity matrix, got {tuple(x.shape)}")
        return _MaxMarginRankFunction.apply(x.float().contiguous(), self.margin, self.lamda1, self.fix_norm)
