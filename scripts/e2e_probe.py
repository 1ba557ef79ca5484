"""Probe: H2D bandwidth of the pinned waveform batch (idle and under a running step), and e2e step time with
synchronous vs pipelined loss reads."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder
from texttoaudiogrounding_b200.models.match import DotProduct
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
from texttoaudiogrounding_b200.train import FusedTrainStep

B, L = 64, 320000
torch.manual_seed(1)
model = BiEncoder(Cnn8Rnn(32000), EmbeddingAgg(5221, 512), DotProduct(), 512).cuda().train()
ts = FusedTrainStep(model)
g = torch.Generator().manual_seed(0)
host = {"waveform": (0.1 * torch.randn(B, L, generator=g)).pin_memory(),
        "waveform_len": torch.full((B,), L, dtype=torch.long).pin_memory(),
        "text": torch.randint(2, 5221, (B, 8), generator=g).pin_memory(),
        "text_len": torch.full((B,), 8, dtype=torch.long).pin_memory(),
        "label": (torch.rand(B, 251, generator=g) > 0.5).float().pin_memory()}
hosts = [host, {k: v.clone().pin_memory() for k, v in host.items()}]
for _ in range(4):
    ts.step(host)
torch.cuda.synchronize()
dst = torch.empty(B, L, device="cuda")
cs = torch.cuda.Stream()
def h2d_ms(n=5, busy=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    if busy:
        for _ in range(n + 1):
            ts.step(None)
    with torch.cuda.stream(cs):
        e0.record()
        for _ in range(n):
            dst.copy_(host["waveform"], non_blocking=True)
        e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
print("H2D 82 MB idle ms", h2d_ms(), " under load ms", h2d_ms(busy=True))
def loop(mode, steps=30):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts.prefetch(hosts[0])
    pending = None
    for i in range(steps):
        if mode == "sync":
            loss = ts.step(hosts[i % 2])
            if i + 1 < steps:
                ts.prefetch(hosts[(i + 1) % 2])
            loss.item()
        else:
            h = ts.step_async(hosts[i % 2])
            if i + 1 < steps:
                ts.prefetch(hosts[(i + 1) % 2])
            if pending is not None:
                pending.result()
            pending = h
    if pending is not None:
        pending.result()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps
for mode in ("sync", "async", "sync", "async"):
    print(mode, "ms/step", round(loop(mode), 3))
