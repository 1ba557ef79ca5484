"""Generate tests/golden/weak_b3_1s.npz by running the UNMODIFIED reference MultiTextBiEncoder
(models/audio_text_model.py:101-229) + ClipBceLoss / ClipFrameBceLoss (losses.py) on a seeded synthetic
weak-phrase batch.  Build container only:   python oracle/make_golden_weak.py
TEST INFRASTRUCTURE ONLY (see oracle/make_golden.py)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import tag_oracle as O  # noqa: E402
from oracle.make_golden import dropout_identity, subsample  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
# embedding sharpening 100 (not the 300 of the strong cases): phrases as short as ONE token have an up to sqrt(8)
# larger mean embedding than the 8-token strong-case phrases; 100 keeps the logits in the same +-10 range
SHARPEN = 100.0
CASE = dict(batch=3, n_samples=32000, n_phrases=5, n_tokens=6, seed=4, data_seed=7)


def main():
    ns = ref_shim.import_reference()
    import models.audio_text_model as atm
    import losses
    torch.manual_seed(0)
    sd = O.synth_state_dict(seed=CASE["seed"], sharpen=SHARPEN, perturb_bn=True)
    batch = O.synth_weak_batch(CASE["batch"], CASE["n_samples"], CASE["n_phrases"], CASE["n_tokens"],
                               seed=CASE["data_seed"])
    out = {"text": batch["text"].numpy(), "text_len": batch["text_len"].numpy(),
           "waveform_len": batch["waveform_len"], "label": batch["label"].numpy(),
           "waveform_checksum": np.array([batch["waveform"].double().sum().item()])}

    def build(pooling):
        m = atm.MultiTextBiEncoder(ns.Cnn8Rnn(32000), ns.EmbeddingAgg(O.VOCAB, 512), ns.DotProduct(), 512,
                                   text_forward_keys=["text"], pooling=pooling)
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        return m

    def fwd(m):
        b = {k: (v.long() if k == "text" else v.float()) if isinstance(v, torch.Tensor) else v for k, v in batch.items()}
        d = {"specaug": False}
        d.update(b)
        return m(d)

    for pooling in ("linear_softmax", "max", "mean", "exp_softmax"):
        m = build(pooling).eval()
        with torch.no_grad():
            o = fwd(m)
        out[f"eval_clip_sim/{pooling}"] = o["clip_sim"].numpy()
        if pooling == "linear_softmax":
            out["eval_frame_sim"] = o["frame_sim"].numpy()
            out["eval_length"] = o["length"].numpy()

    for tag, loss_fn in (("clip", losses.ClipBceLoss()), ("clipframe", losses.ClipFrameBceLoss(frame_weight=0.3))):
        m = build("linear_softmax").train()
        with dropout_identity():
            o = fwd(m)
            o.update({k: (v.float() if isinstance(v, torch.Tensor) else v) for k, v in batch.items() if k != "text"})
            loss = loss_fn(o)
            loss.backward()
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
        out[f"train_loss/{tag}"] = np.array(loss.item())
        out[f"train_total_norm/{tag}"] = np.array(float(total))
        out[f"train_clip_sim/{tag}"] = o["clip_sim"].detach().numpy()
        for n, p in m.named_parameters():
            out[f"grad_norm/{tag}/{n}"] = np.array(p.grad.double().pow(2).sum().sqrt().item())
            out[f"grad_sub/{tag}/{n}"] = subsample(p.grad, 128)
        print(tag, "loss", loss.item(), "norm", float(total))
    fs = torch.as_tensor(out["eval_frame_sim"]).double()
    lg = torch.log(fs / (1 - fs))
    print("logits range", lg.min().item(), lg.max().item())
    np.savez_compressed(os.path.join(OUT, "weak_b3_1s.npz"), **out)


if __name__ == "__main__":
    main()
