#!/usr/bin/env python
"""Where does the bf16 error of the bn0 gradients come from?  Full benchmark size (bs = 64, 10 s, ragged), fp32 CUDA path as
the yardstick (it agrees with the CPU oracle to cosine 0.999+): cosine of the bn0 / block-1 gradients for the bf16 step
with the one-pass Cin = 1 layer on and off, over several weight / data seeds (how much of a difference is noise).  Usage: python scripts/bn0_grad_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import tag_oracle as O  # noqa: E402  (synthetic weights / batch only)
from helpers import build_model  # noqa: E402
from texttoaudiogrounding_b200 import ops  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402

def run(sd, batch, prec, **flags):
    for k, v in flags.items():
        setattr(ops, k, v)
    model = build_model(sd, prec).train()
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
    loss = ts.step(batch).item()
    torch.cuda.synchronize()
    g = {n: p.grad.detach().double().flatten().clone() for n, p in model.named_parameters()}
    del ts, model
    torch.cuda.empty_cache()
    return loss, g


def cos(a, b):
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))


KEYS = ["audio_encoder.bn0.weight", "audio_encoder.bn0.bias", "audio_encoder.conv_block1.conv1.weight",
        "audio_encoder.conv_block1.bn1.weight", "audio_encoder.conv_block1.conv2.weight"]
print("columns:", " | ".join(k.replace("audio_encoder.", "") for k in KEYS), "| min over all parameters")
for sd_seed, data_seed in ((1, 21), (1, 22), (2, 23), (4, 24)):
    sd = O.synth_state_dict(seed=sd_seed, sharpen=30.0, perturb_bn=True)
    batch = O.synth_batch(64, 320000, seed=data_seed, ragged=True)
    l32, g32 = run(sd, batch, "fp32")
    for label, flags in (("ON ", dict(USE_C1_FUSE=True)), ("OFF", dict(USE_C1_FUSE=False))):
        l16, g16 = run(sd, batch, "bf16", **flags)
        row = "  ".join(f"{cos(g16[k], g32[k]):.4f}" for k in KEYS)
        print(f"weights {sd_seed} data {data_seed} one-pass c1 {label}: {row}   min {min(cos(g16[k], g32[k]) for k in g32):.4f}"
              f"   loss {l16:.5f} / {l32:.5f}")
