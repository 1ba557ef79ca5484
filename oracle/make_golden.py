"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported read-only
from /root/reference through oracle/ref_shim.py) on seeded synthetic inputs.

Run in the build container only:   python oracle/make_golden.py
The fixtures are committed; tests compare (a) the oracle restatement and (b) the CUDA
path against them.  TEST INFRASTRUCTURE ONLY.

Fixture contents per case (all float32 unless noted):
  inputs ........ text ids, text_len, waveform_len, label, waveform checksum (inputs are
                  re-synthesised from the seed by oracle.tag_oracle.synth_batch)
  eval forward .. logmel_db (full), per-stage summaries + strided subsamples, audio
                  embedding (full), seq_emb, logits, frame_sim, length
  train step .... (dropout patched to identity; BN in train mode) loss, total grad
                  norm, per-parameter grad norm + strided subsample, post-step BN
                  running stats, per-parameter post-Adam subsample
"""
import contextlib
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import tag_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (batch, n_samples, ragged, state-dict seed)
    "cfg1_b4_2s": dict(batch=4, n_samples=64000, ragged=False, seed=1, data_seed=0),
    "ragged_b4_1s": dict(batch=4, n_samples=32000, ragged=True, seed=2, data_seed=3),
}
SHARPEN = 300.0


def subsample(t: torch.Tensor, n: int = 512) -> np.ndarray:
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].float().numpy().copy()


def summary(t: torch.Tensor) -> np.ndarray:
    t = t.detach().float()
    return np.array([t.mean().item(), t.std().item(), t.abs().max().item(),
                     t.double().pow(2).sum().sqrt().item()], dtype=np.float64)


@contextlib.contextmanager
def dropout_identity():
    import torch.nn.functional as F
    orig = F.dropout
    F.dropout = lambda x, p=0.5, training=True, inplace=False: x
    try:
        yield
    finally:
        F.dropout = orig


def run_case(ns, name, cfg):
    torch.manual_seed(0)
    sd = O.synth_state_dict(seed=cfg["seed"], sharpen=SHARPEN, perturb_bn=True)
    batch = O.synth_batch(cfg["batch"], cfg["n_samples"], seed=cfg["data_seed"],
                          ragged=cfg["ragged"])
    model = ref_shim.build_reference_model(ns, {k: v.clone() for k, v in sd.items()})
    out = {}
    out["text"] = batch["text"].numpy()
    out["text_len"] = batch["text_len"].numpy()
    out["waveform_len"] = batch["waveform_len"]
    out["label"] = batch["label"].numpy()
    out["waveform_checksum"] = np.array([batch["waveform"].double().sum().item(),
                                         batch["waveform"].double().abs().sum().item()])

    # ---- eval forward with hooks on the reference's own modules
    stages = {}
    ae = model.audio_encoder
    hooks = []
    def hook(nm):
        def f(mod, inp, outp):
            stages[nm] = outp.detach() if isinstance(outp, torch.Tensor) else outp[0].detach()
        return f
    for nm in ["db_transform", "bn0", "conv_block1", "conv_block2", "conv_block3",
               "conv_block4", "fc1", "rnn"]:
        hooks.append(getattr(ae, nm).register_forward_hook(hook(nm)))
    model.eval()
    with torch.no_grad():
        o = ref_shim.reference_runner_forward(model, dict(batch), training=False)
    for h in hooks:
        h.remove()
    out["eval_logmel_db"] = stages["db_transform"].numpy()
    for nm in ["bn0", "conv_block1", "conv_block2", "conv_block3", "conv_block4", "fc1"]:
        out[f"eval_{nm}_summary"] = summary(stages[nm])
        out[f"eval_{nm}_sub"] = subsample(stages[nm])
    out["eval_embedding"] = stages["rnn"].numpy()
    out["eval_frame_sim"] = o["frame_sim"].numpy()
    out["eval_length"] = o["length"].numpy()
    fs = o["frame_sim"].double()
    out["eval_logits"] = torch.log(fs / (1 - fs)).float().numpy()

    # ---- one train step, dropout = identity, BN in train mode, clip 1.0, Adam 1e-3
    model.train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    loss_fn = ns.FrameBceLoss()
    with dropout_identity():
        opt.zero_grad()
        o = ref_shim.reference_runner_forward(model, dict(batch), training=True)
        out["train_frame_sim"] = o["frame_sim"].detach().numpy()
        loss = loss_fn(o)
        loss.backward()
    total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    out["train_loss"] = np.array(loss.item())
    out["train_total_norm"] = np.array(float(total_norm))
    names = [n for n, _ in model.named_parameters()]
    out["param_names"] = np.array(names)
    coef = min(1.0, 1.0 / (float(total_norm) + 1e-6))
    for n, p in model.named_parameters():
        g = p.grad / coef          # undo the in-place clip: store raw gradients
        out[f"grad_norm/{n}"] = np.array(g.double().pow(2).sum().sqrt().item())
        out[f"grad_sub/{n}"] = subsample(g, 256)
    opt.step()
    for n, p in model.named_parameters():
        out[f"post_sub/{n}"] = subsample(p, 256)
    for n, b in model.named_buffers():
        if "running_" in n:
            out[f"post_buf/{n}"] = b.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", loss.item(), "norm", float(total_norm),
          "logits range", out["eval_logits"].min(), out["eval_logits"].max())


def init_fixture(ns):
    """Per-tensor checksums of the reference's own initialisation under torch.manual_seed(1):
    pins that the B200 modules draw the same initial weights for the same seed."""
    torch.manual_seed(1)
    model = ns.BiEncoder(ns.Cnn8Rnn(32000), ns.EmbeddingAgg(O.VOCAB, 512), ns.DotProduct(), 512)
    out = {}
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point:
            out["sum/" + k] = np.array(v.double().sum().item())
            out["abs/" + k] = np.array(v.double().abs().sum().item())
    np.savez_compressed(os.path.join(OUT, "init_seed1.npz"), **out)


def main():
    os.makedirs(OUT, exist_ok=True)
    ns = ref_shim.import_reference()
    init_fixture(ns)
    if os.environ.get("TAG_GOLDEN_INIT_ONLY"):
        return
    torch.set_num_threads(os.cpu_count())
    for name, cfg in CASES.items():
        run_case(ns, name, cfg)


if __name__ == "__main__":
    main()
