#!/usr/bin/env python
"""Micro-benchmark of the §8f head paths at the BASELINE sizes (B = 64 clips, T' = 250 frames, D = 512):
  * weak multi-phrase head: 32 phrases per clip (frame_sim [64, 250, 32]) + linear-softmax pooling, fwd + bwd
  * sentence-level alignment: all 64 x 64 pairs, N = 8 tokens, fused score + pooling, fwd + bwd, next to the same math
    through torch ops that materialise the [64, 64, 250, 8] matrix (what the reference executes)
  * attention heads (B = 32, configs[3]): SelfAttention text encoder, CrossAttentionGating + token DotProduct,
    match.CrossAttention, fwd + bwd
CUDA events on the launching stream after warm-up, L2 flushed between iterations; ms per call and algorithmic
GFLOP / MB with the achieved rate.  Usage: python scripts/bench_heads.py [json-out]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texttoaudiogrounding_b200.models import sim_pooling as sp  # noqa: E402
from texttoaudiogrounding_b200.models.align import DotProduct as AlignDot  # noqa: E402
from texttoaudiogrounding_b200.models.cross_encoder import CrossAttentionGating  # noqa: E402
from texttoaudiogrounding_b200.models.match import CrossAttention, DotProduct  # noqa: E402
from texttoaudiogrounding_b200.models.text_encoder import SelfAttention  # noqa: E402
from texttoaudiogrounding_b200.models.utils import pool_with_lens  # noqa: E402

ROWS = []
FLUSH = None


def timeit(fn, n=7):
    """Median ms of ``fn``: warm-up eagerly, capture ONE call into a CUDA graph (these paths are 5-25 short launches, so
    eager timing measures the Python / ctypes launch overhead, not the kernels) and time graph replays."""
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    run, mode = fn, "eager"
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        run, mode = g.replay, "graph"
    except Exception as e:          # noqa: BLE001
        print("graph capture failed, timing eagerly:", type(e).__name__, str(e)[:120])
        torch.cuda.synchronize()
    run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        FLUSH.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    timeit.mode = mode
    return sorted(ts)[len(ts) // 2]


def report(name, ms, gflop=None, mb=None):
    row = {"path": name, "ms": round(ms, 4), "timed": getattr(timeit, "mode", "eager")}
    if gflop is not None:
        row.update({"GFLOP": round(gflop, 2), "TFLOPs": round(gflop / ms, 2)})
    if mb is not None:
        row.update({"MB": round(mb, 1), "GBps": round(mb / ms, 1)})
    ROWS.append(row)
    print(row)


def main():
    torch.manual_seed(0)
    dev = "cuda"
    B, T, D = 64, 250, 512
    audio = (torch.randn(B, T, D, device=dev) * 0.5).requires_grad_(True)
    alen = torch.full((B,), T, dtype=torch.long, device=dev)

    # ---- weak multi-phrase head
    n = 32
    seq = (torch.randn(B, n, D, device=dev) * 0.5).requires_grad_(True)
    dot = DotProduct()
    w = torch.randn(B, n, device=dev)

    def weak():
        audio.grad = seq.grad = None
        clip = pool_with_lens(dot.forward_multi(audio, seq), alen, "linear_softmax")
        (clip * w).sum().backward()
    report("weak head fwd+bwd (64 clips x 32 phrases)", timeit(weak), gflop=3 * 2.0 * B * T * n * D / 1e9,
           mb=(2 * B * T * D * 4 + 2 * B * n * D * 4 + 4 * B * T * n * 4) / 1e6)

    # ---- sentence-level alignment, fused vs materialised
    N = 8
    text = (torch.randn(B, N, D, device=dev) * 0.1).requires_grad_(True)
    tlen = torch.randint(1, N + 1, (B,)).to(dev)
    wa = torch.randn(B, B, device=dev)
    pool = sp.AudioMeanTextMean()
    match = AlignDot(scaled=False)

    def fused():
        audio.grad = text.grad = None
        out = pool({"sim": match(audio, text), "audio_len": alen, "text_len": tlen})
        (out * wa).sum().backward()
    ms_f = timeit(fused)
    report("align fused fwd+bwd (64x64 pairs, N=8)", ms_f, gflop=4 * 2.0 * B * T * B * N * D / 1e9)
    tmask = (torch.arange(N, device=dev)[None, :] < tlen[:, None]).float()
    tl = tlen.float()

    def materialised():           # the reference's op sequence (models/align.py + sim_pooling.AudioMeanTextMean)
        audio.grad = text.grad = None
        score = torch.sigmoid(audio.reshape(-1, D) @ text.reshape(-1, D).t()).clamp(1e-7, 1.0)
        sim = score.reshape(B, T, B, N).transpose(1, 2)                  # [B,B,T,N]
        s = sim.mean(2)                                                  # all clips full length here
        out = (s * tmask[None]).sum(-1) / tl[None]
        (out * wa).sum().backward()
    ms_m = timeit(materialised)
    report("align via torch ops, 4-D matrix materialised (cuBLAS + ATen)", ms_m,
           gflop=3 * 2.0 * B * T * B * N * D / 1e9, mb=B * B * T * N * 4 * 6 / 1e6)
    with torch.no_grad():
        a = pool({"sim": match(audio, text), "audio_len": alen, "text_len": tlen})
        score = torch.sigmoid(audio.reshape(-1, D) @ text.reshape(-1, D).t()).clamp(1e-7, 1.0)
        b = ((score.reshape(B, T, B, N).transpose(1, 2).mean(2)) * tmask[None]).sum(-1) / tl[None]
        print("fused vs torch max abs diff", (a - b).abs().max().item())

    # ---- attention heads, configs[3] batch
    Bc, Nt = 32, 8
    tokens = torch.randint(2, 5221, (Bc, Nt), device=dev)
    tl2 = torch.randint(1, Nt + 1, (Bc,)).to(dev)
    a32 = (torch.randn(Bc, T, D, device=dev) * 0.5).requires_grad_(True)
    al2 = torch.full((Bc,), T, dtype=torch.long, device=dev)
    enc = SelfAttention(5221, D, 8, dropout=0.2).to(dev).train()
    gate = CrossAttentionGating(D).to(dev).train()
    cross = CrossAttention(D, 8, 0.2).to(dev).train()
    tok_dot = DotProduct(text_level="token")
    wf = torch.randn(Bc, T, device=dev)

    def text_enc():
        enc.zero_grad(set_to_none=True)
        o = enc({"text": tokens, "text_len": tl2})
        (o["seq_emb"].sum() + o["token_emb"].sum()).backward()
    report("SelfAttention text encoder fwd+bwd (32 x 8 tokens)", timeit(text_enc))

    def gating():
        enc.zero_grad(set_to_none=True); gate.zero_grad(set_to_none=True); a32.grad = None
        t = enc({"text": tokens, "text_len": tl2})
        o = gate({"audio_emb": a32, "text_emb": t, "audio_len": al2, "text_len": tl2})
        (tok_dot(o) * wf).sum().backward()
    report("SelfAttention + CrossAttentionGating + token DotProduct fwd+bwd (32 x 250 frames)", timeit(gating),
           gflop=3 * 2.0 * Bc * T * D * D * 3 / 1e9)

    def crossattn():
        enc.zero_grad(set_to_none=True); cross.zero_grad(set_to_none=True); a32.grad = None
        t = enc({"text": tokens, "text_len": tl2})
        (cross({"audio_emb": a32, "text_emb": t, "text_len": tl2}) * wf).sum().backward()
    report("SelfAttention + match.CrossAttention fwd+bwd (32 x 250 frames)", timeit(crossattn),
           gflop=3 * 2.0 * Bc * T * D * D * 2 / 1e9)

    if len(sys.argv) > 1:
        json.dump({"device": torch.cuda.get_device_name(0), "rows": ROWS}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
