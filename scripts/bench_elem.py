#!/usr/bin/env python
"""Micro-benchmark of the HBM-bound kernels on the model's shapes (B = 64, 10 s clips, bf16 mode):
ms, algorithmic GB/s and the fraction of the measured HBM peak (MEASURED_PEAKS.json) per launch.
CUDA events on the launching stream, L2 flushed between iterations.
Usage: python scripts/bench_elem.py [json-out]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from texttoaudiogrounding_b200 import engine, ops  # noqa: E402
from texttoaudiogrounding_b200.ops import call  # noqa: E402

B = 64
BLOCKS = [(1001, 64, 64, 2, 2), (500, 32, 128, 2, 2), (250, 16, 256, 1, 2), (250, 8, 512, 1, 2)]   # H, W, C, ph, pw
PEAK = 6553.9
try:
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass


def timeit(fn, n=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


ROWS = []


def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    ROWS.append({"kernel": name, "ms": round(ms, 4), "MB": round(nbytes / 1e6, 1), "GBps": round(gbs, 1),
                 "frac_hbm": round(gbs / PEAK, 3)})
    print(f"{name:46s} {ms:7.3f} ms {nbytes / 1e6:8.1f} MB {gbs:7.0f} GB/s  {100 * gbs / PEAK:5.1f} % of HBM peak")


def main():
    dev = "cuda"
    bf = torch.bfloat16
    f32 = dict(device=dev, dtype=torch.float32)
    # ---- frontend
    L = 320000
    T0 = L // 320 + 1
    wav = 0.1 * torch.randn(B, L, **f32)
    from texttoaudiogrounding_b200.frontend_consts import hann_window, slaney_mel_fbanks
    window = hann_window(1024).to(dev)
    fb = slaney_mel_fbanks(513, 50.0, 14000.0, 64, 32000).to(dev)
    mr = engine.compute_mel_range(fb)
    db = torch.empty(B, T0, 64, **f32)
    st = torch.zeros(128, device=dev, dtype=torch.float64)
    nnz = engine.mel_nnz(mr)
    report("logmel_fwd_v2 (warp per frame pair, fp32 wav)",
           timeit(lambda: call("tag_logmel_fwd_v2", wav, 0, B, L, L, window, fb, mr, nnz, db, st)),
           wav.numel() * 4 + db.numel() * 4)
    wav16 = wav.half()
    report("logmel_fwd_v2 (fp16 wav)",
           timeit(lambda: call("tag_logmel_fwd_v2", wav16, 2, B, L, L, window, fb, mr, nnz, db, st)),
           wav.numel() * 2 + db.numel() * 4)
    report("logmel_fwd (smem Stockham, general filterbank)",
           timeit(lambda: call("tag_logmel_fwd", wav, B, L, L, window, fb, mr, db, st)),
           wav.numel() * 4 + db.numel() * 4)
    x0 = torch.empty(B, T0, 64, device=dev, dtype=bf)
    sc = torch.rand(512, **f32) + 0.5
    sh = torch.randn(512, **f32) * 0.1
    mean = torch.randn(512, **f32) * 0.1
    invstd = torch.rand(512, **f32) + 0.5
    report("scale_shift_act bn0 (f32->bf16)", timeit(lambda: ops.scale_shift_act(db, x0, sc, sh, 64, relu=False)),
           db.numel() * 6)
    # ---- conv1_1
    w1 = torch.randn(64, 9, **f32) * 0.1
    y1 = torch.empty(B, T0, 64, 64, device=dev, dtype=bf)
    st1 = torch.zeros(128, device=dev, dtype=torch.float64)
    report("conv_c1_fwd", timeit(lambda: call("tag_conv_c1_fwd", x0, w1, y1, ops.BF16, st1, B, T0, 64)),
           x0.numel() * 2 + y1.numel() * 2)
    dy1 = torch.randn(B, T0, 64, 64, device=dev).to(bf)
    dw1 = torch.zeros(64, 9, **f32)
    dx0 = torch.empty(B, T0, 64, **f32)
    report("conv_c1_bwd (wgrad+dgrad)", timeit(lambda: call("tag_conv_c1_bwd", dy1, x0, w1, ops.BF16, dw1, dx0, B, T0, 64)),
           dy1.numel() * 2 + x0.numel() * 2 + dx0.numel() * 4)
    del dy1, y1
    red0 = torch.zeros(128, device=dev, dtype=torch.float64)
    report("bn_bwd_reduce_f32 (bn0)", timeit(lambda: call("tag_bn_bwd_reduce_f32", dx0, db, mean, invstd, B * T0, 64, red0)),
           dx0.numel() * 8)
    seed_dev = torch.zeros(1, device=dev, dtype=torch.int64)
    for (H, W, C, ph, pw) in BLOCKS:
        y = torch.randn(B, H, W, C, device=dev).to(bf)
        a = torch.empty_like(y)
        nb = y.numel() * 2
        report(f"scale_shift_act relu {H}x{W}x{C}", timeit(lambda: ops.scale_shift_act(y, a, sc, sh, C, relu=True)), 2 * nb)
        Ho, Wo = H // ph, W // pw
        p = torch.empty(B, Ho, Wo, C, device=dev, dtype=bf)
        report(f"bn_relu_pool_fwd {H}x{W}x{C} pool{ph}{pw}",
               timeit(lambda: call("tag_bn_relu_pool_fwd", y, p, None, ops.BF16, sc, sh, B, H, W, C, ph, pw, 0.2, 123, seed_dev)),
               nb + p.numel() * 2)
        dp = torch.randn(B, Ho, Wo, C, device=dev).to(bf)
        red = torch.zeros(2 * C, device=dev, dtype=torch.float64)
        dy = torch.empty_like(y)
        report(f"bn_relu_pool_bwd reduce {H}x{W}x{C} pool{ph}{pw}",
               timeit(lambda: call("tag_bn_relu_pool_bwd", 0, y, dp, None, ops.BF16, sc, sh, mean, invstd, red, 1, B, H, W, C,
                                   ph, pw, 0.2, 123, seed_dev)), nb + dp.numel() * 2)
        report(f"bn_relu_pool_bwd apply  {H}x{W}x{C} pool{ph}{pw}",
               timeit(lambda: call("tag_bn_relu_pool_bwd", 1, y, dp, dy, ops.BF16, sc, sh, mean, invstd, red, 1, B, H, W, C,
                                   ph, pw, 0.2, 123, seed_dev)), 2 * nb + dp.numel() * 2)
        report(f"bn_relu_bwd apply (no pool) {H}x{W}x{C}",
               timeit(lambda: call("tag_bn_relu_pool_bwd", 1, y, a, dy, ops.BF16, sc, sh, mean, invstd, red, 1, B, H, W, C,
                                   0, 0, 0.0, 0, None)), 3 * nb)
        del y, a, p, dp, dy
    # ---- head-side passes
    rows = B * 250
    x4 = torch.randn(B, 250, 4, 512, device=dev).to(bf)
    m = torch.empty(rows, 512, device=dev, dtype=bf)
    report("freq_mean_fwd", timeit(lambda: call("tag_freq_mean_fwd", x4, m, ops.BF16, rows, 4, 512, 0.5, 77, seed_dev)),
           x4.numel() * 2 + m.numel() * 2)
    dgi = torch.randn(rows, 1536, device=dev).to(bf)
    out = torch.zeros(1536, **f32)
    report("colsum 16000x1536 bf16", timeit(lambda: call("tag_colsum", dgi, ops.BF16, rows, 1536, out)), dgi.numel() * 2)
    n = 8804800
    pbuf, g, mm, vv = (torch.randn(n, **f32) for _ in range(4))
    vv.abs_()
    ss = torch.zeros(1, device=dev, dtype=torch.float64)
    step = torch.zeros(1, device=dev, dtype=torch.int64)
    norm = torch.zeros(1, **f32)
    report("sumsq", timeit(lambda: call("tag_sumsq", g, n, ss)), n * 4)
    report("clip_adam", timeit(lambda: call("tag_clip_adam", pbuf, g, mm, vv, n, ss, step, 1.0, 1.0, 1e-3, None, 0.9, 0.999, 1e-8, norm)),
           n * 4 * 7)
    if len(sys.argv) > 1:
        os.makedirs(os.path.dirname(sys.argv[1]) or ".", exist_ok=True)
        json.dump({"hbm_peak_gbs": PEAK, "rows": ROWS}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
