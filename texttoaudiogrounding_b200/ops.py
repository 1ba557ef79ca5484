"""Thin torch-tensor wrappers over the C ABI (``include/tag_b200.h``).

Each wrapper passes raw device pointers and the current CUDA stream; nothing here computes
on the host and nothing falls back to PyTorch ops.  ``LAUNCHES`` counts the kernels launched
through this module (bench.py reports it as ``gpu_launches``).
"""
from __future__ import annotations

import torch

from . import _lib

F32, BF16 = 0, 1
LAUNCHES = 0
# kernels per C-ABI call (default 1)
_KERNELS_PER_CALL = {"tag_clip_adam": 2, "tag_dot_sigmoid_bwd": 2}


def dt(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


def call(name: str, *args) -> None:
    """Invoke ``name`` with tensors converted to device pointers, on the current stream."""
    global LAUNCHES
    fn = getattr(_lib.lib(), name)
    conv = []
    for a in args:
        if isinstance(a, torch.Tensor):
            if not a.is_cuda:
                raise _lib.TagError(f"{name}: tensor argument is not on a CUDA device "
                                    "(this path has no CPU implementation)")
            conv.append(a.data_ptr())
        else:
            conv.append(a)
    conv.append(torch.cuda.current_stream().cuda_stream)
    _lib.check(fn(*conv), name)
    LAUNCHES += _KERNELS_PER_CALL.get(name, 1)


def conv_fwd(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, taps):
    annotate(f"fwd M={B * H * W} N={Cout} K={taps * Cin}", 2.0 * B * H * W * Cout * taps * Cin)
    call("tag_conv_fwd", x, dt(x), w, y, dt(y), bias, int(relu), stats, B, H, W, Cin, Cout, taps)


def conv_wgrad(dy, x, dw, B, H, W, Cin, Cout, taps, splits):
    annotate(f"wgrad P={B * H * W} Cout={Cout} K={taps * Cin}", 2.0 * B * H * W * Cout * taps * Cin)
    call("tag_conv_wgrad", dy, dt(dy), x, dt(x), dw, B, H, W, Cin, Cout, taps, splits)


def wgrad_splits(P: int, Cin: int, Cout: int, taps: int) -> int:
    bn = 128 if Cin % 128 == 0 else 64
    base = (Cout // 64) * (Cin // bn) * taps
    s = max(1, (148 * 6 + base - 1) // base)
    s = min(s, max(1, P // 256), 4096)
    return s


def bn_finalize(stats, count, C, gamma, beta, rm, rv, momentum, eps, training, update_running,
                scale, shift, mean, invstd):
    call("tag_bn_finalize", stats, float(count), C, gamma, beta, rm, rv, float(momentum), float(eps),
         int(training), int(update_running), scale, shift, mean, invstd)


def scale_shift_act(x, y, scale, shift, C, relu):
    call("tag_scale_shift_act", x, dt(x), y, dt(y), scale, shift, x.numel(), C, int(relu))


# ---------------------------------------------------------------------------------------------
# Live per-kernel timing (bench.py): when PROFILE is a list, every C-ABI call is bracketed by
# CUDA events on the launching stream and recorded as [name, tag, flops, bytes, start, end].
PROFILE = None
_PENDING_META = None


def annotate(tag: str, flops: float = 0.0, nbytes: float = 0.0) -> None:
    """Attach a shape tag and the algorithmic FLOPs / bytes to the next call (profiling only)."""
    global _PENDING_META
    if PROFILE is not None:
        _PENDING_META = (tag, flops, nbytes)


_plain_call = call


def call(name: str, *args) -> None:  # noqa: F811  (profiling wrapper around the plain call)
    global _PENDING_META
    if PROFILE is None:
        return _plain_call(name, *args)
    meta, _PENDING_META = _PENDING_META or ("", 0.0, 0.0), None
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    _plain_call(name, *args)
    e.record()
    PROFILE.append([name, meta[0], meta[1], meta[2], s, e])
