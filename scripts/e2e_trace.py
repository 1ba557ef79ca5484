#!/usr/bin/env python
"""Per-step device timeline of the end-to-end loop against the device-resident loop (bs = 64, 10 s clips, bf16).
  A   resident, one batch          ts.step(None)                       (what bench.py `value` timed up to round 2)
  A2  resident, two batches        ts.step(device batch i % 2)         (inputs in HBM, a device-to-device copy per step)
  D   e2e                          prefetch + step_async(host batch)   (what bench.py `e2e` times)
For every loop: ms/step over the whole loop, the mean busy time of a step (event before -> event after its launches),
the mean idle gap between consecutive steps, and the median SM clock / power during the loop (NVML).
Usage: python scripts/e2e_trace.py [steps]"""
import os
import statistics
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_host_batch  # noqa: E402
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn  # noqa: E402
from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder  # noqa: E402
from texttoaudiogrounding_b200.models.match import DotProduct  # noqa: E402
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402

STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 60
torch.manual_seed(1)
model = BiEncoder(Cnn8Rnn(32000, compute_dtype="bf16"), EmbeddingAgg(5221, 512), DotProduct(), 512).cuda().train()
ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, base_seed=1)


class Nvml:
    def __init__(self):
        import pynvml
        self.n = pynvml
        pynvml.nvmlInit()
        self.h = pynvml.nvmlDeviceGetHandleByIndex(0)
        self.samples, self.run = [], False

    def start(self):
        self.samples, self.run = [], True
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def _loop(self):
        while self.run:
            self.samples.append((self.n.nvmlDeviceGetClockInfo(self.h, self.n.NVML_CLOCK_SM),
                                 self.n.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            time.sleep(0.02)

    def stop(self):
        self.run = False
        self.t.join()
        if not self.samples:
            return 0, 0
        return statistics.median(s[0] for s in self.samples), statistics.median(s[1] for s in self.samples)


nv = Nvml()


def run(label, body):
    """body(i) queues step i; events bracket each step's launches on the current stream"""
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)]
    nv.start()
    t0 = time.perf_counter()
    body(ev)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / STEPS
    mhz, watts = nv.stop()
    total = ev[0][0].elapsed_time(ev[-1][1]) / STEPS
    busy = statistics.mean(a.elapsed_time(b) for a, b in ev)
    gaps = [ev[i][1].elapsed_time(ev[i + 1][0]) for i in range(STEPS - 1)]
    print(f"{label:26s} {total:7.3f} ms/step (events) {wall:7.3f} (wall)  busy {busy:7.3f}  gap mean {statistics.mean(gaps):6.3f} "
          f"max {max(gaps):6.3f}  sm {mhz:.0f} MHz  {watts:.0f} W", flush=True)


def resident_one(ev):
    for a, b in ev:
        a.record()
        ts.step(None)
        b.record()


def make_resident_two(dev):
    def body(ev):
        for i, (a, b) in enumerate(ev):
            a.record()
            ts.step(dev[i % 2])
            b.record()
    return body


def make_e2e(hosts):
    def body(ev):
        ts.prefetch(hosts[0])
        pending = []
        for i, (a, b) in enumerate(ev):
            a.record()
            pending.append(ts.step_async(hosts[i % 2]))
            b.record()
            if i + 1 < STEPS:
                ts.prefetch(hosts[(i + 1) % 2])
            if len(pending) > 2:
                pending.pop(0).result()
        for h in pending:
            h.result()
    return body


hosts = [synth_host_batch(64, 100, True), synth_host_batch(64, 200, True)]
dev = [{k: (v.cuda() if torch.is_tensor(v) else v) for k, v in h.items()} for h in hosts]
for _ in range(3):
    ts.step(hosts[0])
make_e2e(hosts)([(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(STEPS)])
t0 = time.perf_counter()
while time.perf_counter() - t0 < 2.5:            # let the power-capped clocks settle first
    for _ in range(10):
        ts.step(None)
    torch.cuda.synchronize()
for rep in range(3):
    run("A  resident, one batch", resident_one)
    run("A2 resident, two batches", make_resident_two(dev))
    run("D  e2e", make_e2e(hosts))
ts.close()
