"""FrameBceLoss — host-side mirror of reference losses.py:11-35.  Masked mean BCE over the
frames t < length[b]; forward and d(loss)/d(frame_sim) in one kernel (csrc/head.cu)."""
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
