"""Host-side orchestration of the Cnn8Rnn forward / backward over the C-ABI kernels.

This is the kernel schedule behind ``models.audio_encoder.Cnn8Rnn`` (the mirror of reference
models/audio_encoder.py:178-232).  It only allocates device buffers and sequences kernel
launches; all arithmetic is in ``csrc/``.

Weights arrive as an ``EncoderWeights`` bundle of kernel-ready tensors:
  bn[i]   = (gamma, beta, running_mean, running_var)   i = 0 (bn0), 1..8 (block b: bn1, bn2)
  conv[i] = fp32 weights with memory layout [Cout][tap][Cin]   i = 0..7
  fc_w [512,512], fc_b [512], w_ih [1536,512], b_ih [1536], w_hh [2,768,256], b_hh [1536]
Gradients are accumulated (+=) into an identically shaped ``EncoderGrads`` bundle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch

from . import ops
from .ops import call

HOP = 320
N_MELS = 64
CHANNELS = [(1, 64), (64, 128), (128, 256), (256, 512)]
POOLS = [(2, 2), (2, 2), (1, 2), (1, 2)]
BN_EPS = 1e-5
BN_MOMENTUM = 0.1
P_BLOCK = 0.2
P_FC = 0.5


@dataclass
class EncoderWeights:
    window: torch.Tensor
    fb: torch.Tensor
    mel_range: Optional[torch.Tensor]
    bn: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]]
    conv: List[torch.Tensor]
    fc_w: torch.Tensor
    fc_b: torch.Tensor
    w_ih: torch.Tensor
    b_ih: torch.Tensor
    w_hh: torch.Tensor
    b_hh: torch.Tensor
    # optional map of ready-made GEMM operands (FusedTrainStep refreshes them with ONE batched launch per
    # step): ("f", i) / ("t", i) forward / dgrad operand of conv i, "fc", "ih", "fc_t", "ih_t"
    prepared: Optional[dict] = None
    # number of non-zero-support filterbank entries, sum(hi - lo) of mel_range (host int; 0 = unknown)
    mel_nnz: int = 0


@dataclass
class EncoderGrads:
    bn: List[Tuple[torch.Tensor, torch.Tensor]]          # (dgamma, dbeta) x 9
    conv: List[torch.Tensor]
    fc_w: torch.Tensor
    fc_b: torch.Tensor
    w_ih: torch.Tensor
    b_ih: torch.Tensor
    w_hh: torch.Tensor
    b_hh: torch.Tensor


@dataclass
class EncoderCtx:
    B: int = 0
    T0: int = 0
    dtype: torch.dtype = torch.float32
    bn_training: bool = True
    seed: int = 0
    seed_dev: Optional[torch.Tensor] = None
    dropout: bool = False
    db: torch.Tensor = None
    x0: torch.Tensor = None
    bn_aux: list = field(default_factory=list)      # per BN: (scale, shift, mean, invstd)
    y: list = field(default_factory=list)           # raw conv outputs, 8 tensors
    a: list = field(default_factory=list)           # relu(bn1(y1)) per block, 4 tensors
    p: list = field(default_factory=list)           # pooled block outputs, 4 tensors
    dims: list = field(default_factory=list)        # (H, W) at each block's conv resolution
    m: torch.Tensor = None
    f: torch.Tensor = None
    out: torch.Tensor = None
    gates: torch.Tensor = None
    c1_fused: bool = False                          # block 1: y[0] is None, a[0] = relu(bn1(conv1)) in one pass
    pcnt: list = field(default_factory=list)        # per block: open-gate counts of the pooling windows (uint8) or None


class _ZeroArena:
    """The float64 reduction targets of one pass (BatchNorm statistics, BN-backward sums) are slices of one arena
    zeroed by a single fill, instead of one fill kernel per target (20 per train step)."""

    def __init__(self, dev, doubles: int):
        self.buf = torch.zeros(doubles, device=dev, dtype=torch.float64)
        self.off = 0

    def __call__(self, n: int) -> torch.Tensor:
        if self.off + n > self.buf.numel():
            return torch.zeros(n, device=self.buf.device, dtype=torch.float64)
        out = self.buf[self.off:self.off + n]
        self.off += n
        return out


def _seed_for(seed: int, layer: int) -> int:
    return (seed * 1000003 + layer * 7919 + 12345) & 0xFFFFFFFFFFFFFFFF


def _operand(Wt: EncoderWeights, key, make):
    if Wt.prepared is not None and key in Wt.prepared:
        return Wt.prepared[key]
    return make()


LOGMEL_FB_CAP = 2048      # FB_CAP of csrc/frontend.cu


def mel_nnz(mel_range: torch.Tensor) -> int:
    """sum of the filter supports (one host read, done once per model/device)."""
    return int((mel_range[:, 1] - mel_range[:, 0]).clamp_min(0).sum().item())


def compute_mel_range(fb: torch.Tensor) -> torch.Tensor:
    """[64,2] int32 (lo, hi) support of each mel filter, derived from the fb buffer."""
    nz = fb != 0
    any_nz = nz.any(dim=0)
    idx = torch.arange(fb.shape[0], device=fb.device).unsqueeze(1)
    big = fb.shape[0]
    lo = torch.where(nz, idx, torch.full_like(idx, big)).min(dim=0).values
    hi = torch.where(nz, idx + 1, torch.zeros_like(idx)).max(dim=0).values
    lo = torch.where(any_nz, lo, torch.zeros_like(lo))
    return torch.stack([lo, hi], dim=1).to(torch.int32).contiguous()


def encoder_forward(Wt: EncoderWeights, wav: torch.Tensor, *, training: bool, bn_training: bool,
                    dropout: bool, seed: int, dtype: torch.dtype, save: bool,
                    seed_dev: Optional[torch.Tensor] = None,
                    stages: Optional[dict] = None) -> Tuple[torch.Tensor, Optional[EncoderCtx]]:
    """wav [B, L] fp32 (cuda) -> embedding [B, T', 512] fp32.  ``save`` keeps what backward needs."""
    assert wav.is_cuda and wav.dtype in (torch.float32, torch.float16) and wav.dim() == 2
    if wav.stride(1) != 1:
        wav = wav.contiguous()
    dev = wav.device
    B, L = wav.shape
    T0 = L // HOP + 1
    f32 = dict(device=dev, dtype=torch.float32)
    act = dict(device=dev, dtype=dtype)
    ctx = EncoderCtx(B=B, T0=T0, dtype=dtype, bn_training=bn_training, seed=seed, seed_dev=seed_dev,
                     dropout=dropout and training) if save else None
    use_dropout = dropout and training

    zeros64 = _ZeroArena(dev, 4096 if bn_training else 0)

    def bn_aux(C):
        return tuple(torch.empty(C, **f32) for _ in range(4))

    def finalize(stats, count, C, bn_idx, aux):
        g, b, rm, rv = Wt.bn[bn_idx]
        ops.bn_finalize(stats, count, C, g, b, rm, rv, BN_MOMENTUM, BN_EPS, bn_training,
                        bn_training and training, aux[0], aux[1], aux[2], aux[3])

    # ---- log-mel + bn0
    db = torch.empty(B, T0, N_MELS, **f32)
    st = zeros64(2 * N_MELS) if bn_training else None
    ops.annotate(f"logmel B={B} L={L}", 0.0, wav.numel() * wav.element_size() + db.numel() * 4.0)
    if 0 < Wt.mel_nnz <= LOGMEL_FB_CAP:
        # compact (triangular) filterbank: warp-per-frame-pair kernel; fp32 or fp16 waveform (the reference stores
        # waveforms as float16 in its h5 files, utils/data/pack_waveform.py:46-52)
        call("tag_logmel_fwd_v2", wav, 0 if wav.dtype == torch.float32 else 2, B, L, wav.stride(0), Wt.window, Wt.fb,
             Wt.mel_range, Wt.mel_nnz, db, st)
    else:
        if wav.dtype != torch.float32:
            raise ops._lib.TagError("the dense-filterbank log-mel kernel takes fp32 waveforms")
        call("tag_logmel_fwd", wav, B, L, wav.stride(0), Wt.window, Wt.fb, Wt.mel_range, db, st)
    aux0 = bn_aux(N_MELS)
    finalize(st, B * T0, N_MELS, 0, aux0)
    x0 = torch.empty(B, T0, N_MELS, **act)
    ops.scale_shift_act(db, x0, aux0[0], aux0[1], N_MELS, relu=False)
    if stages is not None:
        stages["logmel_db"] = db
        stages["bn0"] = x0
    if save:
        ctx.db, ctx.x0 = db, x0
        ctx.bn_aux.append(aux0)

    # fp32 inference (no statistics, nothing saved): the dense contractions run on the bf16 tensor cores with
    # split-bf16 operands (fp32-accurate to ~2^-16, three times the bf16 cost) instead of CUDA-core FMA
    x3 = dtype == torch.float32 and not save and not bn_training

    def fwd_operand(w, halo_W=None):
        if x3 and ops.x3_eligible(w, halo_W):
            return ops.prep_weight_x3(w, halo_W)
        return ops.prep_weight(w, dtype, halo_W)

    # ---- conv blocks
    x = x0
    H, W = T0, N_MELS
    # the one-pass Cin = 1 layer: its backward (tag_conv_c1_bwd_bn) exists for bf16 and needs the fused reduce of the
    # halo dgrad kernel
    c1_fused = ops.USE_C1_FUSE and (not save or (dtype == torch.bfloat16 and ops.USE_TC and ops.USE_HALO))
    if save:
        ctx.c1_fused = c1_fused
    # the bn2 backward reductions of blocks 1-3 ride in the epilogue of the next block's conv1 dgrad (halo kernel)
    pool_fused = dtype == torch.bfloat16 and ops.USE_TC and ops.USE_HALO and ops.USE_POOL_FUSE
    for blk, ((cin, cout), (ph, pw)) in enumerate(zip(CHANNELS, POOLS)):
        count = B * H * W
        # conv1
        st1 = zeros64(2 * cout) if bn_training else None
        aux1 = bn_aux(cout)
        if cin == 1 and c1_fused:
            # Cin = 1: the batch statistics of conv1's output follow from 54 moments of its 16 MB input, so the layer
            # writes relu(bn1(conv1(x))) in ONE pass; the raw output y1 is never stored (backward recomputes it)
            if bn_training:
                mom = torch.empty(54, device=dev, dtype=torch.float64)
                call("tag_c1_moments", x, ops.dt(x), B, H, W, mom)
                call("tag_c1_stats_from_moments", mom, Wt.conv[0], st1)
            finalize(st1, count, cout, 1, aux1)
            y1 = None
            a1 = torch.empty(B, H, W, cout, **act)
            call("tag_conv_c1_fwd_act", x, Wt.conv[0], aux1[0], aux1[1], a1, ops.dt(a1), B, H, W)
        else:
            y1 = torch.empty(B, H, W, cout, **act)
            if cin == 1:
                call("tag_conv_c1_fwd", x, Wt.conv[0], y1, ops.dt(y1), st1, B, H, W)
            else:
                ops.conv_fwd(x, _operand(Wt, ("f", 2 * blk), lambda: fwd_operand(Wt.conv[2 * blk], W)), y1, None, False, st1, B, H, W, cin, cout, 9)
            finalize(st1, count, cout, 1 + 2 * blk, aux1)
            a1 = torch.empty(B, H, W, cout, **act)
            ops.scale_shift_act(y1, a1, aux1[0], aux1[1], cout, relu=True)
        # conv2
        y2 = torch.empty(B, H, W, cout, **act)
        st2 = zeros64(2 * cout) if bn_training else None
        ops.conv_fwd(a1, _operand(Wt, ("f", 2 * blk + 1), lambda: fwd_operand(Wt.conv[2 * blk + 1], W)), y2, None, False, st2, B, H, W, cout, cout, 9)
        aux2 = bn_aux(cout)
        finalize(st2, count, cout, 2 + 2 * blk, aux2)
        # bn2 + relu + pool + dropout
        Ho, Wo = H // ph, W // pw
        p = torch.empty(B, Ho, Wo, cout, **act)
        # open-gate counts of the pooling windows: with p itself all the backward reductions of bn2 need (engine backward)
        pcnt = torch.empty(B, Ho, Wo, cout, device=dev, dtype=torch.uint8) if (save and pool_fused) else None
        call("tag_bn_relu_pool_fwd", y2, p, pcnt, ops.dt(p), aux2[0], aux2[1], B, H, W, cout, ph, pw,
             P_BLOCK if use_dropout else 0.0, _seed_for(seed, blk), seed_dev)
        if stages is not None:
            stages[f"conv_block{blk + 1}"] = p
        if save:
            ctx.y += [y1, y2]
            ctx.a.append(a1)
            ctx.p.append(p)
            ctx.pcnt.append(pcnt)
            ctx.bn_aux += [aux1, aux2]
            ctx.dims.append((H, W))
        x = p
        H, W = Ho, Wo

    # ---- mean over frequency (+dropout 0.5), fc1 + relu, GRU
    Tp, Wf, C = H, W, CHANNELS[-1][1]
    rows = B * Tp
    m = torch.empty(rows, C, **act)
    call("tag_freq_mean_fwd", x, m, ops.dt(m), rows, Wf, C, P_FC if use_dropout else 0.0,
         _seed_for(seed, 4), seed_dev)
    f = torch.empty(rows, 512, **act)
    ops.conv_fwd(m, _operand(Wt, "fc", lambda: fwd_operand(Wt.fc_w)), f, Wt.fc_b, True, None, 1, rows, 1, C, 512, 1)
    gi = torch.empty(rows, 1536, **f32)
    ops.conv_fwd(f, _operand(Wt, "ih", lambda: fwd_operand(Wt.w_ih)), gi, Wt.b_ih, False, None, 1, rows, 1, 512, 1536, 1)
    out = torch.empty(B, Tp, 512, **f32)
    gates = torch.empty(B, Tp, 2, 4, 256, **f32) if save else None
    gru_fwd = "tag_gru_fwd_bf16" if (dtype == torch.bfloat16 and ops.USE_TC) else "tag_gru_fwd"
    call(gru_fwd, gi, Wt.w_hh, Wt.b_hh, out, gates, B, Tp)
    if stages is not None:
        stages["fc1"] = f.view(B, Tp, 512)
        stages["rnn"] = out
    if save:
        ctx.m, ctx.f, ctx.out, ctx.gates = m, f, out, gates
    return out, ctx


class _SideStream:
    """Weight-gradient GEMMs are off the critical path of the backward pass (nothing but the optimizer reads them),
    so they can run on a second stream next to the HBM-bound elementwise passes of the main stream.  ``run(fn, *t)``
    forks after the work already queued on the main stream, runs ``fn`` on the side stream and keeps the tensors
    it reads alive until ``join()`` (the caching allocator must not hand their memory to a later main-stream
    allocation while the side stream still reads it)."""

    def __init__(self, stream: Optional[torch.cuda.Stream]):
        self.stream = stream
        self.keep = []

    def run(self, fn, *tensors):
        if self.stream is None:
            fn()
            return
        ev = torch.cuda.Event()
        ev.record()
        self.stream.wait_event(ev)
        self.keep.extend(tensors)
        with torch.cuda.stream(self.stream):
            fn()

    def join(self):
        if self.stream is not None:
            ev = torch.cuda.Event()
            ev.record(self.stream)
            torch.cuda.current_stream().wait_event(ev)
        self.keep.clear()


def encoder_backward(Wt: EncoderWeights, ctx: EncoderCtx, d_emb: torch.Tensor, G: EncoderGrads,
                     side_stream: Optional[torch.cuda.Stream] = None, on_block_done=None) -> None:
    """Backward of encoder_forward: accumulates parameter gradients into ``G``.  ``on_block_done(blk)`` is called once the
    kernels producing every gradient of conv block ``blk`` (3 .. 0) and of everything after it are queued — the hook the
    data-parallel step uses to start the all-reduce of the finished part of the bucket under the remaining backward."""
    side = _SideStream(side_stream)
    dev = d_emb.device
    B, dtype = ctx.B, ctx.dtype
    f32 = dict(device=dev, dtype=torch.float32)
    act = dict(device=dev, dtype=dtype)
    Tp = ctx.out.shape[1]
    rows = B * Tp
    d_emb = d_emb.contiguous()
    p_blk = P_BLOCK if ctx.dropout else 0.0
    p_fc = P_FC if ctx.dropout else 0.0
    zeros64 = _ZeroArena(dev, 8192)

    # ---- GRU
    tc = dtype == torch.bfloat16 and ops.USE_TC
    gdt = dict(device=dev, dtype=torch.bfloat16 if tc else torch.float32)   # the bf16 GRU emits GEMM operands
    dgi = torch.empty(rows, 1536, **gdt)
    dgh = torch.empty(2, rows, 768, **gdt)
    hprev = torch.empty(2, rows, 256, **gdt)
    call("tag_gru_bwd_bf16" if tc else "tag_gru_bwd", d_emb, ctx.out, ctx.gates, Wt.w_hh, dgi, dgh, hprev, B, Tp)
    call("tag_colsum", dgi, ops.dt(dgi), rows, 1536, G.b_ih)
    for d in range(2):
        call("tag_colsum", dgh[d], ops.dt(dgh), rows, 768, G.b_hh[d * 768:(d + 1) * 768])
        ops.conv_wgrad(dgh[d], hprev[d], G.w_hh[d], 1, rows, 1, 256, 768, 1,
                       ops.wgrad_splits(rows, 256, 768, 1))
    dgi_op = dgi
    ops.conv_wgrad(dgi_op, ctx.f, G.w_ih, 1, rows, 1, 512, 1536, 1, ops.wgrad_splits(rows, 512, 1536, 1))
    w_ih_t = _operand(Wt, "ih_t", lambda: ops.prep_weight_t(Wt.w_ih, 1536, 512, 1, dtype))
    df = torch.empty(rows, 512, **f32)
    ops.conv_fwd(dgi_op, w_ih_t, df, None, False, None, 1, rows, 1, 1536, 512, 1)

    # ---- fc1
    dpre = torch.empty(rows, 512, **act)
    call("tag_relu_bwd", df, ops.F32, ctx.f, ops.dt(ctx.f), dpre, ops.dt(dpre), df.numel())
    call("tag_colsum", dpre, ops.dt(dpre), rows, 512, G.fc_b)
    ops.conv_wgrad(dpre, ctx.m, G.fc_w, 1, rows, 1, 512, 512, 1, ops.wgrad_splits(rows, 512, 512, 1))
    fc_t = _operand(Wt, "fc_t", lambda: ops.prep_weight_t(Wt.fc_w, 512, 512, 1, dtype))
    dm = torch.empty(rows, 512, **act)
    ops.conv_fwd(dpre, fc_t, dm, None, False, None, 1, rows, 1, 512, 512, 1)

    # ---- frequency mean (+dropout)
    Wf = ctx.p[3].shape[2]
    dp = torch.empty_like(ctx.p[3])
    red_next = None
    if ctx.pcnt[3] is not None and dp.dtype == torch.bfloat16:
        # block 4's bn2 backward sums ride along (activation domain), as blocks 1-3 get theirs from the next conv1 dgrad
        red_next = zeros64(2 * CHANNELS[3][1])
        call("tag_freq_mean_bwd", dm, ops.dt(dm), dp, ops.dt(dp), rows, Wf, 512, p_fc,
             _seed_for(ctx.seed, 4), ctx.seed_dev, ctx.p[3], ctx.pcnt[3], red_next)
    else:
        call("tag_freq_mean_bwd", dm, ops.dt(dm), dp, ops.dt(dp), rows, Wf, 512, p_fc,
             _seed_for(ctx.seed, 4), ctx.seed_dev, None, None, None)

    # ---- conv blocks, last to first
    bn_tr = int(ctx.bn_training)
    for blk in range(3, -1, -1):
        cin, cout = CHANNELS[blk]
        ph, pw = POOLS[blk]
        H, W = ctx.dims[blk]
        y1, y2 = ctx.y[2 * blk], ctx.y[2 * blk + 1]
        a1 = ctx.a[blk]
        aux1, aux2 = ctx.bn_aux[1 + 2 * blk], ctx.bn_aux[2 + 2 * blk]
        P = B * H * W
        # bn2 + relu + pool + dropout
        seed = _seed_for(ctx.seed, blk)
        if red_next is not None:
            # the two reductions came with dp (epilogue of the previous iteration's conv1 dgrad), in the activation domain
            red, red_next = red_next, None
            dg, dbt = G.bn[2 + 2 * blk]
            call("tag_bn_red_act_to_xhat", red, Wt.bn[2 + 2 * blk][0], Wt.bn[2 + 2 * blk][1], cout,
                 0.25 / (1.0 - p_blk) if p_blk > 0.0 else 0.25,        # the codes carry 4 x the weight; dropout keep scale
                 dg, dbt)
        else:
            red = zeros64(2 * cout)
            call("tag_bn_relu_pool_bwd", 0, y2, dp, None, ops.dt(y2), aux2[0], aux2[1], aux2[2], aux2[3], red,
                 bn_tr, B, H, W, cout, ph, pw, p_blk, seed, ctx.seed_dev)
            dg, dbt = G.bn[2 + 2 * blk]
            call("tag_bn_param_grads", red, cout, dg, dbt)
        dy2 = torch.empty_like(y2)
        call("tag_bn_relu_pool_bwd", 1, y2, dp, dy2, ops.dt(y2), aux2[0], aux2[1], aux2[2], aux2[3], red,
             bn_tr, B, H, W, cout, ph, pw, p_blk, seed, ctx.seed_dev)
        # conv2
        side.run(lambda: ops.conv_wgrad(dy2, a1, G.conv[2 * blk + 1], B, H, W, cout, cout, 9,
                                        ops.wgrad_splits(P, cout, cout, 9)), dy2, a1)
        w2t = _operand(Wt, ("t", 2 * blk + 1), lambda: ops.prep_weight_t(Wt.conv[2 * blk + 1], cout, cout, 9, dtype, W))
        da1 = torch.empty_like(a1)
        red1 = zeros64(2 * cout)
        if cin == 1 and ctx.c1_fused:
            # the BN input y1 was never stored; the fused reduce works on the saved activation a1 anyway
            ops.conv_fwd(dy2, w2t, da1, None, False, red1, B, H, W, cout, cout, 9, bn_fuse=a1)
            dg, dbt = G.bn[1]
            call("tag_bn_red_act_to_xhat", red1, Wt.bn[1][0], Wt.bn[1][1], cout, 1.0, dg, dbt)
            del dy2
            # conv1 backward with the BatchNorm backward applied on the fly (y1 recomputed from x0)
            dx0 = torch.empty(B, H, W, **f32)
            call("tag_conv_c1_bwd_bn", da1, ctx.x0, Wt.conv[0], aux1[0], aux1[2], aux1[3], red1, bn_tr, G.conv[0], dx0,
                 B, H, W)
            del da1
            red0 = zeros64(2 * N_MELS)
            aux0 = ctx.bn_aux[0]
            call("tag_bn_bwd_reduce_f32", dx0, ctx.db, aux0[2], aux0[3], B * H, N_MELS, red0)
            dg, dbt = G.bn[0]
            call("tag_bn_param_grads", red0, N_MELS, dg, dbt)
            if on_block_done is not None:
                on_block_done(blk)
            continue
        if ops.can_fuse_bn_bwd(w2t, a1):
            # dgrad with the ReLU gate and the BN-backward reductions of bn1 fused into its epilogue (activation domain)
            ops.conv_fwd(dy2, w2t, da1, None, False, red1, B, H, W, cout, cout, 9, bn_fuse=a1)
            dg, dbt = G.bn[1 + 2 * blk]
            call("tag_bn_red_act_to_xhat", red1, Wt.bn[1 + 2 * blk][0], Wt.bn[1 + 2 * blk][1], cout, 1.0, dg, dbt)
        else:
            ops.conv_fwd(dy2, w2t, da1, None, False, None, B, H, W, cout, cout, 9)
            call("tag_bn_relu_pool_bwd", 0, y1, da1, None, ops.dt(y1), aux1[0], aux1[1], aux1[2], aux1[3], red1,
                 bn_tr, B, H, W, cout, 0, 0, 0.0, 0, None)
            dg, dbt = G.bn[1 + 2 * blk]
            call("tag_bn_param_grads", red1, cout, dg, dbt)
        del dy2
        # bn1 + relu
        dy1 = torch.empty_like(y1)
        call("tag_bn_relu_pool_bwd", 1, y1, da1, dy1, ops.dt(y1), aux1[0], aux1[1], aux1[2], aux1[3], red1,
             bn_tr, B, H, W, cout, 0, 0, 0.0, 0, None)
        del da1
        # conv1
        if cin == 1:
            dx0 = torch.empty(B, H, W, **f32)
            call("tag_conv_c1_bwd", dy1, ctx.x0, Wt.conv[0], ops.dt(dy1), G.conv[0], dx0, B, H, W)
            red0 = zeros64(2 * N_MELS)
            aux0 = ctx.bn_aux[0]
            call("tag_bn_bwd_reduce_f32", dx0, ctx.db, aux0[2], aux0[3], B * H, N_MELS, red0)
            dg, dbt = G.bn[0]
            call("tag_bn_param_grads", red0, N_MELS, dg, dbt)
        else:
            x_in = ctx.p[blk - 1]
            side.run(lambda: ops.conv_wgrad(dy1, x_in, G.conv[2 * blk], B, H, W, cin, cout, 9,
                                            ops.wgrad_splits(P, cin, cout, 9)), dy1, x_in)
            w1t = _operand(Wt, ("t", 2 * blk), lambda: ops.prep_weight_t(Wt.conv[2 * blk], cout, cin, 9, dtype, W))
            dp = torch.empty_like(x_in)
            pc = ctx.pcnt[blk - 1]
            if pc is not None and ops.can_fuse_bn_bwd(w1t, x_in):
                # x_in is the pooled output of block blk - 1: its bn2 backward reductions ride in this dgrad's epilogue
                red_next = zeros64(2 * cin)
                ops.conv_fwd(dy1, w1t, dp, None, False, red_next, B, H, W, cout, cin, 9, bn_fuse=(x_in, pc))
            else:
                ops.conv_fwd(dy1, w1t, dp, None, False, None, B, H, W, cout, cin, 9)
        del dy1
        if on_block_done is not None:
            side.join()              # the hook may read the weight gradients queued on the side stream so far
            on_block_done(blk)
    side.join()
