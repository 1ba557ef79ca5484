import os, sys, time
import torch
sys.path.insert(0, "/root/repo")
from bench import synth_host_batch
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder
from texttoaudiogrounding_b200.models.match import DotProduct
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
from texttoaudiogrounding_b200.train import FusedTrainStep
STEPS = 40
torch.manual_seed(1)
model = BiEncoder(Cnn8Rnn(32000, compute_dtype="bf16"), EmbeddingAgg(5221, 512), DotProduct(), 512).cuda().train()
ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, base_seed=1)
hosts = [synth_host_batch(64, 100, False)]
for _ in range(4):
    ts.step(hosts[0])
def timed(fn, label):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(f"{label:50s} {e0.elapsed_time(e1) / STEPS:7.3f} ms/step")
slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(4)]
def v_copy_nowait():
    for i in range(STEPS):
        loss = ts.step(None)
        slots[i % 4].copy_(loss, non_blocking=True)
def v_event_wait_only():
    evs = []
    for i in range(STEPS):
        ts.step(None)
        ev = torch.cuda.Event(); ev.record(); evs.append(ev)
        if i >= 1: evs[i - 1].synchronize()
def v_copy_side_stream():
    side = torch.cuda.Stream()
    evs = []
    for i in range(STEPS):
        loss = ts.step(None)
        done = torch.cuda.Event(); done.record()
        with torch.cuda.stream(side):
            side.wait_event(done)
            slots[i % 4].copy_(loss, non_blocking=True)
            ev = torch.cuda.Event(); ev.record()
        evs.append(ev)
        if i >= 1: evs[i - 1].synchronize()
def v_async():
    pending = None
    for _ in range(STEPS):
        h = ts.step_async(None)
        if pending is not None: pending.result()
        pending = h
    pending.result()
def v_lag2():
    hs = []
    for i in range(STEPS):
        hs.append(ts.step_async(None))
        if i >= 2: hs[i - 2].result()
    hs[-1].result()
timed(lambda: [ts.step(None) for _ in range(STEPS)], "A resident (queue runs ahead)")
timed(v_copy_nowait, "copy loss D2H each step, never wait")
timed(v_event_wait_only, "no copy; host waits for step i-1 each step")
timed(v_copy_side_stream, "copy on a side stream; host waits for step i-1")
timed(v_async, "step_async + result() of step i-1 (bench e2e)")
timed(v_lag2, "step_async + result() of step i-2")
timed(lambda: [ts.step(None) for _ in range(STEPS)], "A resident again")
