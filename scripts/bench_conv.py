#!/usr/bin/env python
"""Micro-benchmark of the dense tcgen05 kernels on the model's layer shapes (B = 64, 10 s clips).
Usage: python scripts/bench_conv.py [fwd|dgrad|wgrad] — prints ms and TFLOP/s per layer (CUDA events, L2 flushed)."""
import sys
import torch
sys.path.insert(0, ".")
from texttoaudiogrounding_b200 import ops

LAYERS = [  # H, W, Cin, Cout
    (1001, 64, 64, 64), (500, 32, 64, 128), (500, 32, 128, 128), (250, 16, 128, 256),
    (250, 16, 256, 256), (250, 8, 256, 512), (250, 8, 512, 512)]
B = 64


def timeit(fn, n=5):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


def main(which):
    for (H, W, Cin, Cout) in LAYERS:
        x = torch.randn(B, H, W, Cin, device="cuda").bfloat16()
        w32 = torch.randn(Cout, 3, 3, Cin, device="cuda") * 0.05
        flops = 2.0 * B * H * W * Cout * 9 * Cin
        if which == "fwd":
            w = ops.prep_weight(w32, torch.bfloat16, W)
            y = torch.empty(B, H, W, Cout, device="cuda", dtype=torch.bfloat16)
            st = torch.zeros(2 * Cout, device="cuda", dtype=torch.float64)
            for label, stats in (("stats", st), ("nostats", None)):
                ms = timeit(lambda: ops.conv_fwd(x, w, y, None, False, stats, B, H, W, Cin, Cout, 9))
                print(f"fwd {H}x{W} {Cin}->{Cout} {label}: {ms:.3f} ms {flops / ms / 1e9:.0f} TFLOP/s")
        elif which == "dgrad":
            # the halo kernel as dgrad: input dy [.., Cout], output [.., Cin]; with the fused ReLU + BN-backward reduce
            # when the layer is a block's conv2 (Cin == Cout)
            dy = torch.randn(B, H, W, Cout, device="cuda").bfloat16()
            wt = ops.prep_weight_t(w32, Cout, Cin, 9, torch.bfloat16, W)
            dx = torch.empty(B, H, W, Cin, device="cuda", dtype=torch.bfloat16)
            red = torch.zeros(2 * Cin, device="cuda", dtype=torch.float64)
            fuse = x if Cin == Cout else None          # the saved activation of the layer (any bf16 tensor of that shape here)
            ms = timeit(lambda: ops.conv_fwd(dy, wt, dx, None, False, red if fuse is not None else None, B, H, W, Cout, Cin, 9,
                                             bn_fuse=fuse))
            print(f"dgrad {H}x{W} {Cout}->{Cin} {"bn_fuse" if fuse is not None else "plain"}: {ms:.3f} ms {flops / ms / 1e9:.0f} TFLOP/s")
        else:
            dy = torch.randn(B, H, W, Cout, device="cuda").bfloat16()
            dw = torch.zeros(Cout, 3, 3, Cin, device="cuda")
            ms = timeit(lambda: ops.conv_wgrad(dy, x, dw, B, H, W, Cin, Cout, 9, 1))
            print(f"wgrad {H}x{W} {Cin}->{Cout}: {ms:.3f} ms {flops / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fwd")
