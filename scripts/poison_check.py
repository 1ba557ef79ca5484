#!/usr/bin/env python
"""Does any kernel of the train step read memory it (or an earlier kernel) did not write?  torch.empty / empty_like are patched
to hand out NaN-filled (floating point) or 0x7F-filled (integer) buffers; one eager train step then must still produce a finite
loss and finite gradients, identical to the unpoisoned run.
Usage: python scripts/poison_check.py [fp32|bf16] [seconds] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scripts.dp_check import build, small_batch  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
seconds = int(sys.argv[2]) if len(sys.argv) > 2 else 1
B = int(sys.argv[3]) if len(sys.argv) > 3 else 4

_empty, _empty_like = torch.empty, torch.empty_like
POISON = {"on": False}


def _poison(t):
    if POISON["on"] and t.is_cuda and t.numel() > 0:
        if t.dtype.is_floating_point:
            t.fill_(float("nan"))
        elif t.dtype in (torch.uint8, torch.int8):
            t.fill_(127)
        else:
            t.fill_(0x7F7F7F7F)
    return t


torch.empty = lambda *a, **k: _poison(_empty(*a, **k))
torch.empty_like = lambda *a, **k: _poison(_empty_like(*a, **k))

model = build(7, prec)
model.audio_encoder.dropout_enabled = False
ts = FusedTrainStep(model, lr=0.0, max_grad_norm=1e9, base_seed=1, use_graph=False)
batch = small_batch(B, 50, seconds=seconds)
out = {}
for mode in ("clean", "poisoned", "poisoned"):
    POISON["on"] = mode == "poisoned"
    ts.step(batch)
    torch.cuda.synchronize()
    g = ts.flat_g.clone()
    key = mode if mode not in out else mode + "2"
    out[key] = (float(ts.loss_out), g, ts.sim.clone())
    print(f"{key:10s} loss {float(ts.loss_out):.9f}  |g| {float(g.double().norm()):.9e}  non-finite grads {int((~torch.isfinite(g)).sum())}  "
          f"non-finite sim {int((~torch.isfinite(ts.sim)).sum())}")
POISON["on"] = False
names = {id(p): n for n, p in model.named_parameters()}
ref = out["clean"][1]
for key in ("poisoned", "poisoned2"):
    g = out[key][1]
    bad = []
    for p, (off, k) in zip(ts._params, ts._views):
        a, b = ref[off:off + k], g[off:off + k]
        d = float((a - b).norm() / a.norm().clamp_min(1e-30)) if torch.isfinite(b).all() else float("inf")
        if d > 1e-5:
            bad.append((d, names[id(p)]))
    print(key, "parameters whose gradient differs from the clean run by > 1e-5:", sorted(bad, reverse=True)[:10] or "none")
