#!/usr/bin/env python
"""Is the encoder forward bit-reproducible?  The same batch N times through engine.encoder_forward (train-mode BatchNorm, dropout
off); every saved stage tensor of run i is compared bit for bit with run 0 and the first stage that differs is reported.
Usage: python scripts/fwd_determinism.py [bf16|fp32] [seconds] [batch] [runs]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.dp_check import build, small_batch  # noqa: E402
from texttoaudiogrounding_b200 import engine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
RUNS = int(sys.argv[4]) if len(sys.argv) > 4 else 6
model = build(7, prec)
enc = model.audio_encoder
Wt = enc._weights()
wav = small_batch(B, 50, seconds=seconds)["waveform"].cuda()


def stages(ctx, emb):
    out = [("db (log-mel)", ctx.db), ("x0 (bn0)", ctx.x0)]
    for i, aux in enumerate(ctx.bn_aux):
        for j, nm in enumerate(("scale", "shift", "mean", "invstd")):
            out.append((f"bn{i}.{nm}", aux[j]))
    for i, t in enumerate(ctx.y):
        out.append((f"y[{i}] (conv out)", t))
    for i, t in enumerate(ctx.a):
        out.append((f"a[{i}] (relu bn1)", t))
    for i, t in enumerate(ctx.p):
        out.append((f"p[{i}] (pooled)", t))
    out += [("m (freq mean)", ctx.m), ("f (fc1)", ctx.f), ("gates", ctx.gates), ("out (GRU)", ctx.out), ("emb", emb)]
    return [(n, t.clone()) for n, t in out if t is not None]


ref = None
junk = []
for r in range(RUNS):
    emb, ctx = engine.encoder_forward(Wt, wav, training=True, bn_training=True, dropout=False, seed=1,
                                      dtype=enc.compute_dtype, save=True, seed_dev=None)
    torch.cuda.synchronize()
    cur = stages(ctx, emb)
    if ref is None:
        ref = cur
    else:
        diffs = [(n, float((a.double() - b.double()).abs().max())) for (n, a), (_, b) in zip(cur, ref) if not torch.equal(a, b)]
        print(f"run {r}: " + ("bit-identical to run 0" if not diffs else
                              f"{len(diffs)} stages differ, first: " + ", ".join(f"{n} (max abs {d:.3e})" for n, d in diffs[:4])))
    junk.append(torch.randn(1 << 20, device="cuda") * r)          # perturb the allocator state between runs
    del ctx, emb
