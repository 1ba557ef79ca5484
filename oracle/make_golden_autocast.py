"""Generate tests/golden/autocast_b64_10s.npz: what the UNMODIFIED reference does at BASELINE.json's full size (bs = 64,
10 s clips, ragged) in fp32 and under torch.autocast(bfloat16) — the yardstick for this repository's bf16 mode where
the north_star's absolute 1e-2 bar cannot hold for any bf16 implementation (x300-sharpened weights: logits beyond +-10).
Contents: frame_sim of both runs (eval, sd seed 3, sharpen 300), and for the train step (sd seed 1, sharpen 30, dropout
off) loss, total norm and the per-parameter cosine between the autocast and the fp32 gradients.
Build container only (CPU autocast):   python oracle/make_golden_autocast.py      TEST INFRASTRUCTURE ONLY."""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import tag_oracle as O  # noqa: E402
from oracle.make_golden import dropout_identity  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
B, L = 64, 320000


def inputs(batch):
    return {"specaug": False, "waveform": batch["waveform"], "waveform_len": batch["waveform_len"],
            "text": batch["text"].long(), "text_len": batch["text_len"]}


def main():
    ns = ref_shim.import_reference()
    batch = O.synth_batch(B, L, seed=21, ragged=True)
    out = {}
    sd = O.synth_state_dict(seed=3, sharpen=300.0, perturb_bn=True)
    m = ref_shim.build_reference_model(ns, {k: v.clone() for k, v in sd.items()}).eval()
    with torch.no_grad():
        f32 = m(inputs(batch))["frame_sim"]
        with torch.autocast("cpu", dtype=torch.bfloat16):
            f16 = m(inputs(batch))["frame_sim"].float()
    out["eval_frame_sim_fp32"] = f32.numpy()
    out["eval_frame_sim_autocast"] = f16.numpy()
    print("eval x300: autocast vs fp32 max abs", (f16 - f32).abs().max().item())

    sd = O.synth_state_dict(seed=1, sharpen=30.0, perturb_bn=True)
    grads = {}
    for mode in ("fp32", "autocast"):
        t0 = time.time()
        m = ref_shim.build_reference_model(ns, {k: v.clone() for k, v in sd.items()}).train()
        with dropout_identity():
            if mode == "autocast":
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    o = m(inputs(batch))
                o["frame_sim"] = o["frame_sim"].float()
            else:
                o = m(inputs(batch))
            T = min(o["frame_sim"].shape[1], batch["label"].shape[1])
            o.update({"frame_sim": o["frame_sim"][..., :T], "label": batch["label"][..., :T].float(),
                      "length": torch.clamp(o["length"], 1, T)})
            loss = ns.FrameBceLoss()(o)
            loss.backward()
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters()))
        grads[mode] = {n: p.grad.detach().double().clone() for n, p in m.named_parameters()}
        out[f"train_loss_{mode}"] = np.array(loss.item())
        out[f"train_total_norm_{mode}"] = np.array(float(total))
        print(mode, "loss", loss.item(), "norm", float(total), f"{time.time() - t0:.1f}s")
    for n in grads["fp32"]:
        a, b = grads["fp32"][n].flatten(), grads["autocast"][n].flatten()
        c = float((a * b).sum() / (a.norm() * b.norm() + 1e-300))
        out[f"grad_cosine/{n}"] = np.array(c)
        out[f"grad_norm_fp32/{n}"] = np.array(float(a.norm()))
    worst = sorted((float(out[k]), k) for k in out if k.startswith("grad_cosine/"))[:6]
    print("lowest autocast-vs-fp32 gradient cosines:", worst)
    np.savez_compressed(os.path.join(OUT, "autocast_b64_10s.npz"), **out)


if __name__ == "__main__":
    main()
