#!/usr/bin/env python
"""The small dense layers around the GRU (M = B * T' = 16 000 rows): fc1, the GRU input projection, their dgrads and
weight gradients, timed in isolation (CUDA events, L2 flushed).  Usage: python scripts/bench_small_gemm.py"""
import sys

import torch

sys.path.insert(0, ".")
from texttoaudiogrounding_b200 import ops  # noqa: E402


def timeit(fn, n=7):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


M = 16000
bf = torch.bfloat16
for name, K, N, out_dt, bias, relu in (("fc1 fwd", 512, 512, bf, True, True), ("W_ih fwd (gi fp32)", 512, 1536, torch.float32, True, False),
                                       ("W_ih dgrad (fp32 out)", 1536, 512, torch.float32, False, False),
                                       ("fc1 dgrad", 512, 512, bf, False, False)):
    x = torch.randn(M, K, device="cuda").to(bf)
    w = (torch.randn(N, K, device="cuda") * 0.05).to(bf)
    y = torch.empty(M, N, device="cuda", dtype=out_dt)
    b = torch.randn(N, device="cuda") if bias else None
    ms = timeit(lambda: ops.conv_fwd(x, w, y, b, relu, None, 1, M, 1, K, N, 1))
    flops = 2.0 * M * K * N
    mb = (x.numel() * 2 + y.numel() * y.element_size()) / 1e6
    print(f"{name:24s} M={M} K={K} N={N}: {ms * 1e3:7.1f} us  {flops / ms / 1e9:6.0f} TFLOP/s  {mb / ms / 1e3:6.2f} TB/s of in+out")
for name, K, N in (("W_hh wgrad", 256, 768), ("W_ih wgrad", 512, 1536), ("fc1 wgrad", 512, 512)):
    dy = torch.randn(M, N, device="cuda").to(bf)
    x = torch.randn(M, K, device="cuda").to(bf)
    dw = torch.zeros(N, K, device="cuda")
    ms = timeit(lambda: ops.conv_wgrad(dy, x, dw, 1, M, 1, K, N, 1, ops.wgrad_splits(M, K, N, 1)))
    print(f"{name:24s} P={M} Cin={K} Cout={N}: {ms * 1e3:7.1f} us  {2.0 * M * K * N / ms / 1e9:6.0f} TFLOP/s")
