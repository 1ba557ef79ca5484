#!/usr/bin/env python
"""Per-parameter agreement of the bf16 tcgen05 train step with the fp32 path at the full benchmark size (bs = 64, 10 s
clips, ragged lengths): loss, total gradient norm, cosine of every parameter gradient (the data behind
tests/test_gpu_fullsize.py).  Usage: python scripts/fullsize_grad_agreement.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tag_oracle as O  # noqa: E402  (synthetic weights / batch only)
from helpers import build_model, cosine, sub  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402

sd = O.synth_state_dict(seed=1, sharpen=30.0, perturb_bn=True)
batch = O.synth_batch(64, 320000, seed=21, ragged=True)
res = {}
for prec in ("fp32", "bf16", "bf16"):
    model = build_model(sd, prec).train()
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
    loss = ts.step(batch).item()
    torch.cuda.synchronize()
    res.setdefault(prec, []).append((loss, ts.norm_out.item(), {n: sub(p.grad, 2048) for n, p in model.named_parameters()}))
    del ts, model
    torch.cuda.empty_cache()
(l32, n32, g32) = res["fp32"][0]
for i, (l16, n16, g16) in enumerate(res["bf16"]):
    cos = sorted((cosine(g16[n], g32[n]), n) for n in g32)
    print(f"bf16 run {i}: loss {l16:.6f} vs {l32:.6f}  norm {n16:.4f} vs {n32:.4f}")
    for c, n in cos[:6]:
        print(f"   {c:.4f}  {n}")
a, b = res["bf16"][0][2], res["bf16"][1][2]
print("run-to-run min cosine (atomics order):", min(cosine(a[n], b[n]) for n in a))
