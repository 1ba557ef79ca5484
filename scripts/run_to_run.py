#!/usr/bin/env python
"""Run-to-run spread of the gradients: the same batch twice through the eager train step (dropout off, lr = 0), per parameter.
Usage: python scripts/run_to_run.py [fp32|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.dp_check import build, small_batch  # noqa: E402
from texttoaudiogrounding_b200 import ops  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp32"
model = build(7, prec)
model.audio_encoder.dropout_enabled = False
if os.environ.get("DET") == "1":
    ops.set_deterministic_wgrad(True)
ts = FusedTrainStep(model, lr=0.0, max_grad_norm=1e9, base_seed=1, use_graph=False)
batch = small_batch(4, 50, seconds=1)
grads = []
p0 = ts.flat_p.clone()
for i in range(4):
    ts.step(batch)
    torch.cuda.synchronize()
    grads.append(ts.flat_g.clone())
    print(f"run {i}: loss {float(ts.loss_out):.9f}  |g| {float(ts.flat_g.double().norm()):.9e}  params moved by "
          f"{float((ts.flat_p - p0).abs().max()):.3e}  sim checksum {float(ts.sim.double().sum()):.9f}")
names = {id(p): n for n, p in model.named_parameters()}
print(f"{prec}: whole bucket, run 1 vs 0: {float((grads[1] - grads[0]).norm() / grads[0].norm()):.3e}   run 2 vs 0: "
      f"{float((grads[2] - grads[0]).norm() / grads[0].norm()):.3e}")
rows = []
CMP = int(os.environ.get("CMP", "1"))
print(f"per parameter, run {CMP} vs run 0:")
for p, (off, k) in zip(ts._params, ts._views):
    a, b = grads[0][off:off + k], grads[CMP][off:off + k]
    rows.append((float((a - b).norm() / a.norm().clamp_min(1e-30)), float((a - b).abs().max()), float(a.norm()), names[id(p)]))
for r in sorted(rows, reverse=True)[:40]:
    print(f"  rel {r[0]:.3e}  max abs {r[1]:.3e}  |g| {r[2]:.3e}  {r[3]}")
