"""Thin torch-tensor wrappers over the C ABI (``include/tag_b200.h``).

Each wrapper passes raw device pointers and the current CUDA stream; nothing here computes
on the host and nothing falls back to PyTorch ops.  ``LAUNCHES`` counts the kernels launched
through this module (bench.py reports it as ``gpu_launches``).
"""
from __future__ import annotations

import os

import torch

from . import _lib

F32, BF16 = 0, 1
LAUNCHES = 0
# kernels per C-ABI call (default 1)
_KERNELS_PER_CALL = {"tag_clip_adam": 2, "tag_dot_sigmoid_bwd": 2, "tag_multi_dot_sigmoid_bwd": 2}


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


def set_sm_reserve(sms: int) -> None:
    """Persistent tensor-core kernels launched from now on leave ``sms`` SMs free (0 = use all); see tag_set_sm_reserve."""
    _lib.check(_lib.lib().tag_set_sm_reserve(int(sms)), "tag_set_sm_reserve")


_SPLITK_WS = None


def set_deterministic_wgrad(on: bool, device=None, megabytes: int = 64) -> None:
    """Deterministic (two-pass) split-K for the tensor-core weight gradients: every split writes its own slab of a
    workspace, a reduce adds the slabs in split order (tag_set_splitk_workspace).  Off = fp32 atomics (default)."""
    global _SPLITK_WS
    if on:
        _SPLITK_WS = torch.empty(megabytes << 20, device=device or "cuda", dtype=torch.uint8)
        _lib.check(_lib.lib().tag_set_splitk_workspace(_SPLITK_WS.data_ptr(), _SPLITK_WS.numel()), "tag_set_splitk_workspace")
    else:
        _lib.check(_lib.lib().tag_set_splitk_workspace(None, 0), "tag_set_splitk_workspace")
        _SPLITK_WS = None


USE_TC = os.environ.get("TAG_B200_NO_TC", "0") != "1"
USE_HALO = os.environ.get("TAG_B200_NO_HALO", "0") != "1"
USE_C1_FUSE = os.environ.get("TAG_B200_NO_C1_FUSE", "0") != "1"      # conv_block1.conv1 + bn1 + relu in one pass
USE_POOL_FUSE = os.environ.get("TAG_B200_NO_POOL_FUSE", "0") != "1"  # bn2 backward reductions in the next dgrad's epilogue


def tc_eligible(W: int, Cin: int, Cout: int) -> bool:
    return USE_TC and Cin % 64 == 0 and Cout % 64 == 0 and W in (1, 2, 4, 8, 16, 32, 64)


def conv_fwd(x, w, y, bias, relu, stats, B, H, W, Cin, Cout, taps, bn_fuse=None):
    """Dispatch on the weight dtype: bf16 weights -> tcgen05 kernel, fp32 weights -> SIMT kernel.
    ``bn_fuse`` = the saved activation relu(bn(.)) of the layer this dgrad differentiates through: output gated by
    a > 0, ``stats`` = (sum g | sum g * a) — the fused ReLU + BN-backward reduce of the halo kernel; or the triple
    pair (pooled output, open-gate codes) when the dgrad's output is the gradient of a pooled block output."""
    if getattr(w, "_tag_x3", False):
        # fp32 activations x split-bf16 weights: fp32-accurate product on the bf16 tensor cores (csrc/split.cu)
        if x.dtype != torch.float32 or y.dtype != torch.float32 or stats is not None or bn_fuse is not None:
            raise _lib.TagError("split-bf16 conv: fp32 activations in and out, no fused statistics")
        xs = torch.empty(B * H * W, 3 * Cin, device=x.device, dtype=torch.bfloat16)
        call("tag_split_bf16x3", x, xs, B * H * W, Cin, 0, 0)
        annotate(f"fwd M={B * H * W} N={Cout} K={taps * 3 * Cin}", 2.0 * B * H * W * Cout * taps * 3 * Cin)
        if taps == 9:
            call("tag_conv_tc_fwd_halo", xs, w, y, dt(y), None, B, H, W, 3 * Cin, Cout, None, None)
        else:
            call("tag_conv_tc_fwd", xs, w, y, dt(y), bias, int(relu), None, B, H, W, 3 * Cin, Cout, taps)
        return
    annotate(f"fwd M={B * H * W} N={Cout} K={taps * Cin}", 2.0 * B * H * W * Cout * taps * Cin)
    if w.dtype == torch.bfloat16:
        if x.dtype != torch.bfloat16 or not tc_eligible(W, Cin, Cout):
            raise _lib.TagError("tag_conv_tc_fwd: needs bf16 activations and Cin, Cout multiples of 64")
        if getattr(w, "_tag_tapmajor", False):
            if taps != 9 or bias is not None or relu or W % 8 != 0:
                raise _lib.TagError("tap-major weights are only valid for the 3x3 halo kernel")
            act, cnt = bn_fuse if isinstance(bn_fuse, tuple) else (bn_fuse, None)
            call("tag_conv_tc_fwd_halo", x, w, y, dt(y), stats, B, H, W, Cin, Cout, act, cnt)
            return
        if bn_fuse is not None:
            raise _lib.TagError("bn_fuse needs the halo kernel")
        call("tag_conv_tc_fwd", x, w, y, dt(y), bias, int(relu), stats, B, H, W, Cin, Cout, taps)
    else:
        if bn_fuse is not None:
            raise _lib.TagError("bn_fuse needs the halo kernel")
        call("tag_conv_fwd", x, dt(x), w, y, dt(y), bias, int(relu), stats, B, H, W, Cin, Cout, taps)


def can_fuse_bn_bwd(w, act) -> bool:
    return bool(getattr(w, "_tag_tapmajor", False)) and act.dtype == torch.bfloat16


def conv_wgrad(dy, x, dw, B, H, W, Cin, Cout, taps, splits, tc=None):
    annotate(f"wgrad P={B * H * W} Cout={Cout} K={taps * Cin}", 2.0 * B * H * W * Cout * taps * Cin)
    if tc is None:
        tc = (dy.dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and tc_eligible(W, Cin, Cout)
              and (taps == 9 or Cin >= 128))
    if tc and USE_HALO and taps == 9 and Cin == 64 and W % 8 == 0:
        call("tag_conv_tc_wgrad64", dy, x, dw, B, H, W, Cout, max(1, 148 // (Cout // 64)))
    elif tc:
        call("tag_conv_tc_wgrad", dy, x, dw, B, H, W, Cin, Cout, taps, tc_wgrad_splits(B, H, W, Cin, Cout, taps))
    else:
        call("tag_conv_wgrad", dy, dt(dy), x, dt(x), dw, B, H, W, Cin, Cout, taps, splits)


def tc_wgrad_splits(B, H, W, Cin, Cout, taps) -> int:
    bn = 256 if Cout % 256 == 0 else (128 if Cout % 128 == 0 else 64)
    pix = 64 if taps == 1 else {256: 64, 128: 128, 64: 256}[bn]
    thk = max(1, pix // min(W, pix))
    k_tiles = B * ((H + thk - 1) // thk)
    m_blocks = taps * (Cin // 128) if Cin >= 128 else (taps + 1) // 2
    base = m_blocks * (Cout // bn)
    # one CTA (198 KB of smem) is resident per SM, so the grid runs in waves of 148: pick the split count that
    # fills two waves from BELOW (297 CTAs would run three waves for the work of two)
    s = max(1, (148 * 2) // base)
    return max(1, min(s, k_tiles // 8 if k_tiles >= 8 else 1))


def _halo_ok(halo_W) -> bool:
    return USE_HALO and halo_W is not None and halo_W % 8 == 0


def prep_weight(w: torch.Tensor, dtype: torch.dtype, halo_W=None) -> torch.Tensor:
    """fp32 master weight -> GEMM operand of the compute dtype.  bf16: [Cout][taps*Cin] (same layout)
    or, for a 3x3 conv on a feature map of width ``halo_W`` (multiple of 8), the tap-major
    [9][Cout][Cin] operand of the halo kernel (marked with ``_tag_tapmajor``)."""
    if dtype != torch.bfloat16 or not USE_TC:
        return w
    wb = torch.empty(w.shape, device=w.device, dtype=torch.bfloat16)
    if w.dim() == 4 and _halo_ok(halo_W):
        call("tag_weight_prep_tapmajor_bf16", w, wb, w.shape[0], w.shape[3], 0)
        wb._tag_tapmajor = True
    else:
        call("tag_cast_f32_to_bf16", w, wb, w.numel())
    return wb


def x3_eligible(w: torch.Tensor, halo_W) -> bool:
    """3x3 conv on a width that suits the halo kernel, or a linear layer; channel counts multiples of 64."""
    if not USE_TC or os.environ.get("TAG_B200_FP32_X3", "1") == "0":
        return False
    if w.dim() == 4:
        return _halo_ok(halo_W) and w.shape[0] % 64 == 0 and w.shape[3] % 64 == 0
    return w.dim() == 2 and w.shape[0] % 64 == 0 and w.shape[1] % 64 == 0


def prep_weight_x3(w: torch.Tensor, halo_W=None) -> torch.Tensor:
    """fp32 master weight -> split-bf16 forward operand ([9][Cout][3Cin] tap-major for a 3x3 conv, [Cout][3Cin] for a
    linear layer), marked ``_tag_x3`` so conv_fwd splits the fp32 activations to match."""
    if w.dim() == 4:
        Co, Ci = w.shape[0], w.shape[3]
        out = torch.empty(9 * Co * 3 * Ci, device=w.device, dtype=torch.bfloat16)
        call("tag_weight_prep_tapmajor_x3", w, out, Co, Ci)
    else:
        Co, Ci = w.shape
        out = torch.empty(Co * 3 * Ci, device=w.device, dtype=torch.bfloat16)
        call("tag_split_bf16x3", w, out, Co, Ci, 1, 0)
    out._tag_x3 = True
    return out


def prep_weight_t(w: torch.Tensor, Co: int, Ci: int, taps: int, dtype: torch.dtype, halo_W=None) -> torch.Tensor:
    """fp32 master [Co][taps][Ci] -> flipped + transposed [Ci][taps][Co] dgrad operand."""
    if dtype == torch.bfloat16 and USE_TC:
        wt = torch.empty(Ci * taps * Co, device=w.device, dtype=torch.bfloat16)
        if taps == 9 and _halo_ok(halo_W):
            call("tag_weight_prep_tapmajor_bf16", w, wt, Co, Ci, 1)
            wt._tag_tapmajor = True
        else:
            call("tag_weight_flip_transpose_bf16", w, wt, Co, Ci, taps)
    else:
        wt = torch.empty(Ci * taps * Co, device=w.device, dtype=torch.float32)
        call("tag_weight_flip_transpose", w, wt, Co, Ci, taps)
    return wt


class WeightPrep:
    """Persistent bf16 operand buffers for a fixed set of fp32 master weights + one batched prep launch.
    ``add(src, mode, rows, cols, taps)`` registers an operand and returns its bf16 tensor;
    ``run()`` refreshes all of them (modes: 0 cast, 1 tap-major fwd, 2 tap-major dgrad, 3 transpose)."""
    ELEMS = 2048

    def __init__(self, device):
        self.device = device
        self.entries = []
        self.table = None
        self.blocks = 0

    def add(self, src: torch.Tensor, mode: int, rows: int, cols: int, taps: int) -> torch.Tensor:
        dst = torch.empty(src.numel(), device=self.device, dtype=torch.bfloat16)
        if mode in (1, 2):
            dst._tag_tapmajor = True
        self.entries.append((src, dst, mode, rows, cols, taps))
        self.table = None
        return dst

    def key(self):
        return tuple(e[0].data_ptr() for e in self.entries)

    def run(self):
        if self.table is None:
            rows, blk = [], 0
            for src, dst, mode, r, c, t in self.entries:
                rows.append([src.data_ptr(), dst.data_ptr(), (r << 32) | c, (t << 32) | mode, blk])
                blk += (src.numel() + self.ELEMS - 1) // self.ELEMS
            self.table = torch.tensor(rows, dtype=torch.int64).to(self.device)
            self.blocks = blk
        call("tag_weight_prep_batch", self.table, len(self.entries), self.blocks)


def to_bf16(x: torch.Tensor) -> torch.Tensor:
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    call("tag_cast_f32_to_bf16", x, y, x.numel())
    return y


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
