#!/usr/bin/env python
"""Data-parallel correctness of FusedTrainStep on real GPUs (NCCL), one process per GPU:
  1. replicas that start from different seeds are identical after construction (rank 0's state is broadcast);
  2. after K steps on DIFFERENT batches per rank every rank holds bit-identical parameters and Adam state;
  3. the reduced gradient bucket equals the sum of the ranks' local gradients (fp32 mode, dropout off, eager): every rank
     recomputes all ranks' local gradients on one GPU without a collective and compares with what the all-reduce left in
     flat_g (relative error of the whole bucket and of its late / early parts, which travel as two calls with
     TAG_B200_AR_OVERLAP=1).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 scripts/dp_check.py
Prints one line per check on rank 0 and exits non-zero on a mismatch."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import synth_host_batch  # noqa: E402
from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn  # noqa: E402
from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder  # noqa: E402
from texttoaudiogrounding_b200.models.match import DotProduct  # noqa: E402
from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg  # noqa: E402
from texttoaudiogrounding_b200.train import FusedTrainStep  # noqa: E402


def build(seed, precision):
    torch.manual_seed(seed)
    return BiEncoder(Cnn8Rnn(32000, compute_dtype=precision), EmbeddingAgg(5221, 512), DotProduct(), 512).cuda().train()


def small_batch(B, seed, seconds=2):
    b = synth_host_batch(B, seed, False)
    n = 32000 * seconds
    b["waveform"] = b["waveform"][:, :n].contiguous()
    b["waveform_len"] = torch.full((B,), n, dtype=torch.long)
    b["label"] = b["label"][:, :51].contiguous()
    return b


def same_everywhere(t, world):
    chk = torch.stack([t.double().sum(), t.double().abs().sum(), t.double().pow(2).sum()])
    got = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(got, chk)
    return all(torch.equal(got[0], g) for g in got)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    if os.environ.get("TAG_B200_AR_OVERLAP", "0") == "1":
        os.environ.setdefault("NCCL_MAX_CTAS", "8")
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    ok = True

    def report(name, good, extra=""):
        nonlocal ok
        flags = torch.tensor([1.0 if good else 0.0], device="cuda")
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        ok = ok and bool(flags.item())
        if rank == 0:
            print(f"{'ok  ' if flags.item() else 'FAIL'} {name} {extra}", flush=True)

    # ---- 1 + 2: bf16 production step, CUDA graphs, different seeds and different batches per rank
    model = build(100 + rank, "bf16")
    ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, base_seed=1)
    report("replicas identical after construction", same_everywhere(ts.flat_p, world))
    for step in range(5):
        ts.step(small_batch(8, 1000 * rank + step))
    torch.cuda.synchronize()
    report("parameters identical after 5 steps (bf16, graphs, dropout)", same_everywhere(ts.flat_p, world))
    report("Adam moments identical", same_everywhere(ts.flat_m, world) and same_everywhere(ts.flat_v, world))
    bn = model.audio_encoder.conv_block2.bn1.running_mean
    report("BatchNorm running statistics stay per rank (no SyncBN, as the reference)",
           world == 1 or not same_everywhere(bn, world))
    ts.close()

    # ---- 3: the reduced bucket = sum of the local gradients (fp32 mode, no dropout, eager)
    model = build(7, "fp32")
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, lr=0.0, max_grad_norm=1e9, base_seed=1, use_graph=False)
    batches = [small_batch(4, 50 + r, seconds=1) for r in range(world)]
    ts.step(batches[rank])
    reduced = ts.flat_g.clone()
    w, ts.world = ts.world, 1                       # local gradients of every rank's batch, no collective
    local_sum = torch.zeros_like(reduced)
    def rel(a, b):
        return float((a - b).norm() / b.norm().clamp_min(1e-30))
    last_local = None
    for r in range(world):
        ts.step(batches[r])
        local_sum += ts.flat_g
        last_local = ts.flat_g.clone()
    ts.step(batches[world - 1])                    # the same batch once more: run-to-run spread of the atomics-based sums
    again = rel(ts.flat_g, last_local)
    ts.world = w
    split = ts._ar_split

    e_all, e_late, e_early = rel(reduced, local_sum), rel(reduced[:split], local_sum[:split]), rel(reduced[split:], local_sum[split:])
    report("all-reduced bucket = sum of local gradients", max(e_all, e_late, e_early) < max(1e-3, 10 * again),
           f"(rel. error whole {e_all:.2e}, head [{split} floats] {e_late:.2e}, tail {e_early:.2e}; the same batch run twice "
           f"on one GPU differs by {again:.2e}: the order of the atomics-based sums; "
           f"overlap={'on' if ts._overlap_ar else 'off'})")
    ts.close()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
