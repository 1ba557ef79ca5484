"""Host-side helpers mirroring reference models/utils.py (init_weights 5-19, generate_length_mask 22-30,
the *_with_lens pooling functions 33-95).  The poolings of a CUDA [B, T, n] (or [B, T]) tensor run in
csrc/multitext.cu (one thread per (clip, phrase) scanning the valid frames) with an analytic backward."""
import torch
import torch.nn as nn

from ..ops import call

POOL_MODES = {"linear_softmax": 0, "max": 1, "mean": 2, "exp_softmax": 3}


def init_weights(m):
    if isinstance(m, (nn.Conv2d, nn.Conv1d)):
        nn.init.kaiming_normal_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.BatchNorm2d):
        nn.init.constant_(m.weight, 1)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Linear):
        nn.init.kaiming_uniform_(m.weight)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.Embedding):
        nn.init.kaiming_uniform_(m.weight)


def generate_length_mask(lens, max_length=None):
    lens = torch.as_tensor(lens)
    if max_length is None:
        max_length = int(lens.max().item())
    idxs = torch.arange(max_length, device=lens.device).unsqueeze(0)
    return idxs < lens.view(-1, 1)


def lens_to_device(lens, device) -> torch.Tensor:
    """list / numpy / int or float tensor (the reference's Runner.forward hands text_len over as
    float32 on device, run_strong.py:94-99) -> int64 tensor on ``device``."""
    t = torch.as_tensor(lens)
    return t.to(device=device, dtype=torch.long)


class _PoolWithLens(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sim, length, mode):
        B, T, n = sim.shape
        clip = torch.empty(B, n, device=sim.device, dtype=torch.float32)
        call("tag_pool_with_lens_fwd", sim, length, mode, clip, B, T, n)
        ctx.save_for_backward(sim, clip, length)
        ctx.mode = mode
        return clip

    @staticmethod
    def backward(ctx, d_clip):
        sim, clip, length = ctx.saved_tensors
        B, T, n = sim.shape
        d_sim = torch.empty_like(sim)
        call("tag_pool_with_lens_bwd", d_clip.contiguous(), sim, clip, length, ctx.mode, d_sim, B, T, n)
        return d_sim, None, None


def pool_with_lens(features: torch.Tensor, lens, pooling: str) -> torch.Tensor:
    """features [N, T] or [N, T, n] (CUDA) pooled over the frames t < lens[i]."""
    if pooling not in POOL_MODES:
        raise Exception(f"Unsupported pooling {pooling}")
    if not features.is_cuda:
        raise RuntimeError("*_with_lens (B200) needs CUDA tensors: there is no CPU fallback")
    squeeze = features.ndim == 2
    f = features.unsqueeze(-1) if squeeze else features
    if f.ndim != 3:
        raise NotImplementedError("*_with_lens is implemented for [N, T] and [N, T, n] tensors")
    out = _PoolWithLens.apply(f.float().contiguous(), lens_to_device(lens, f.device).contiguous(), POOL_MODES[pooling])
    return out.squeeze(-1) if squeeze else out


def linear_softmax_with_lens(features, lens):
    return pool_with_lens(features, lens, "linear_softmax")


def max_with_lens(features, lens):
    return pool_with_lens(features, lens, "max")


def mean_with_lens(features, lens):
    return pool_with_lens(features, lens, "mean")


def exp_softmax_with_lens(features, lens):
    return pool_with_lens(features, lens, "exp_softmax")
