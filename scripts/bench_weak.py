#!/usr/bin/env python
"""Weakly supervised multi-phrase train step (MultiTextBiEncoder + ClipBceLoss + Adam, reference
python_scripts/training/run_weak_phrase.py) on one B200: bs = 64 clips x 32 phrases x 8 tokens, 10 s clips, bf16,
WeakFusedTrainStep with CUDA-graph replay, inputs resident in HBM.  Usage: bench_weak.py [json-out]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn  # noqa: E402
from texttoaudiogrounding_b200.models.audio_text_model import MultiTextBiEncoder  # noqa: E402
from texttoaudiogrounding_b200.models.match import DotProduct  # noqa: E402
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg  # noqa: E402
from texttoaudiogrounding_b200.train import WeakFusedTrainStep  # noqa: E402

B, L, n, N = 64, 320000, 32, 8
torch.manual_seed(1)
model = MultiTextBiEncoder(Cnn8Rnn(32000), EmbeddingAgg(5221, 512), DotProduct(), 512, text_forward_keys=["text"],
                           pooling="linear_softmax").cuda().train()
ts = WeakFusedTrainStep(model)
g = torch.Generator().manual_seed(0)
batch = {"waveform": (0.1 * torch.randn(B, L, generator=g)).pin_memory(),
         "waveform_len": torch.full((B,), L, dtype=torch.long),
         "text": torch.randint(2, 5221, (B, n, N), generator=g), "text_len": torch.full((B, n), N, dtype=torch.long),
         "label": (torch.rand(B, n, generator=g) > 0.5).float()}
for _ in range(4):
    ts.step(batch)
torch.cuda.synchronize()
steps = 30
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    ts.step(None)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
row = {"workload": f"weak multi-phrase train step, {B} clips x {n} phrases, 10 s clips, bf16", "ms_per_step": round(ms, 3),
       "clips_per_s": round(B / ms * 1e3, 1), "clip_phrase_pairs_per_s": round(B * n / ms * 1e3, 1),
       "loss": float(ts.loss_out.item())}
print(row)
if len(sys.argv) > 1:
    json.dump({"device": torch.cuda.get_device_name(0), "rows": [row]}, open(sys.argv[1], "w"), indent=1)
