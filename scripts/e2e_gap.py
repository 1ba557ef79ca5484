#!/usr/bin/env python
"""Where does the end-to-end loop lose time against the device-resident step?  bs = 64, 10 s clips, bf16 train step.
  A  resident            ts.step(None), inputs in HBM
  B  + loss read         step_async(None) with the pipelined loss read
  C  + dummy H2D         B plus an 82 MB pinned->device copy per step on a side stream that nobody reads (L2 / HBM interference)
  D  e2e fp32 waveforms  prefetch + step_async(batch)  (what bench.py times)
  E  e2e fp16 waveforms  the same with float16 host waveforms (the reference's h5 storage type)
Usage: python scripts/e2e_gap.py [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_host_batch  # noqa: E402
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn  # noqa: E402
from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder  # noqa: E402
from texttoaudiogrounding_b200.models.match import DotProduct  # noqa: E402
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 40
torch.manual_seed(1)
model = BiEncoder(Cnn8Rnn(32000, compute_dtype="bf16"), EmbeddingAgg(5221, 512), DotProduct(), 512).cuda().train()
ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, base_seed=1)


def timed(fn, label):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / STEPS
    print(f"{label:28s} {e0.elapsed_time(e1) / STEPS:7.3f} ms/step (events)  {wall:7.3f} ms/step (wall)")


def e2e(hosts):
    ts.prefetch(hosts[0])
    pending = None
    for i in range(STEPS):
        h = ts.step_async(hosts[i % 2])
        if i + 1 < STEPS:
            ts.prefetch(hosts[(i + 1) % 2])
        if pending is not None:
            pending.result()
        pending = h
    pending.result()


def loss_only():
    pending = None
    for _ in range(STEPS):
        h = ts.step_async(None)
        if pending is not None:
            pending.result()
        pending = h
    pending.result()


side = torch.cuda.Stream()
scratch = torch.empty(64, 320000, device="cuda")


def dummy_h2d(host_wav):
    pending = None
    for _ in range(STEPS):
        h = ts.step_async(None)
        with torch.cuda.stream(side):
            scratch.copy_(host_wav, non_blocking=True)
        if pending is not None:
            pending.result()
        pending = h
    pending.result()


hosts32 = [synth_host_batch(64, 100, False), synth_host_batch(64, 200, False)]
hosts16 = [synth_host_batch(64, 100, True), synth_host_batch(64, 200, True)]
for h in (hosts16, hosts32):
    for _ in range(3):
        ts.step(h[0])
t0 = time.perf_counter()
while time.perf_counter() - t0 < 2.5:            # let the power-capped clocks settle first
    for _ in range(10):
        ts.step(None)
    torch.cuda.synchronize()
for rep in range(3):
    ts.step(hosts32[0])
    timed(lambda: [ts.step(None) for _ in range(STEPS)], "A resident")
    timed(loss_only, "B + pipelined loss read")
    timed(lambda: dummy_h2d(hosts32[0]["waveform"]), "C + dummy 82 MB H2D")
    timed(lambda: e2e(hosts32), "D e2e fp32 waveforms")
    timed(lambda: e2e(hosts16), "E e2e fp16 waveforms")
