"""Generate tests/golden/attn_b3_1s.npz by running the UNMODIFIED reference BiEncoder (models/audio_text_model.py:16-98)
with the attention-type components of BASELINE.json configs[3]: text_encoder.SelfAttention
(models/text_encoder.py:240-268), match.CrossAttention (models/match.py:63-88), cross_encoder.CrossAttentionGating
(models/cross_encoder.py:60-79) + match.DotProduct(text_level="token").  The attention modules are built with
dropout 0 (their RNG streams cannot be shared with another implementation); the audio encoder's dropout is patched
to identity as in the other fixtures.  Build container only:   python oracle/make_golden_attn.py
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
SHARPEN = 20.0
CASE = dict(batch=3, n_samples=32000, n_tokens=7, seed=6, data_seed=11, attn_seed=13)


def state(variant):
    return O.attn_case_state(variant, CASE["seed"], CASE["attn_seed"], SHARPEN)


def main():
    ns = ref_shim.import_reference()
    import models.text_encoder as te
    import models.match as match
    import models.cross_encoder as ce
    import losses
    torch.manual_seed(0)
    batch = O.synth_batch(CASE["batch"], CASE["n_samples"], CASE["n_tokens"], seed=CASE["data_seed"], ragged=True)
    out = {"text": batch["text"].numpy(), "text_len": batch["text_len"].numpy(),
           "waveform_checksum": np.array([batch["waveform"].double().sum().item()])}

    def build(variant):
        text_encoder = te.SelfAttention(O.VOCAB, 512, O.HEADS, dropout=0.0)
        if variant == "crossattn":
            m = ns.BiEncoder(ns.Cnn8Rnn(32000), text_encoder, match.CrossAttention(512, O.HEADS, 0.0), 512)
        else:
            m = ns.BiEncoder(ns.Cnn8Rnn(32000), text_encoder, match.DotProduct(text_level="token"), 512,
                             cross_encoder=ce.CrossAttentionGating(512))
        m.load_state_dict({k: v.clone() for k, v in state(variant).items()}, strict=True)
        return m

    def inputs():
        return {"specaug": False, "waveform": batch["waveform"], "waveform_len": batch["waveform_len"],
                "text": batch["text"].long(), "text_len": batch["text_len"]}

    loss_fn = losses.FrameBceLoss()
    for variant in ("crossattn", "gating"):
        m = build(variant).eval()
        with torch.no_grad():
            o = m(inputs())
            t = m.text_encoder(inputs())
        out[f"eval_frame_sim/{variant}"] = o["frame_sim"].numpy()
        out[f"eval_length/{variant}"] = o["length"].numpy()
        out[f"eval_seq_emb/{variant}"] = t["seq_emb"].numpy()
        out[f"eval_token_emb/{variant}"] = t["token_emb"].numpy()
        m = build(variant).train()
        with dropout_identity():
            o = m(inputs())
            T = o["frame_sim"].shape[1]
            o["label"] = batch["label"][:, :T]
            o["length"] = torch.as_tensor(o["length"]).clamp(1, T)
            loss = loss_fn(o)
            loss.backward()
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
        out[f"train_loss/{variant}"] = np.array(loss.item())
        out[f"train_total_norm/{variant}"] = np.array(float(total))
        out[f"train_frame_sim/{variant}"] = o["frame_sim"].detach().numpy()
        for n, p in m.named_parameters():
            out[f"grad_norm/{variant}/{n}"] = np.array(p.grad.double().pow(2).sum().sqrt().item())
            out[f"grad_sub/{variant}/{n}"] = subsample(p.grad, 128)
        fs = o["frame_sim"].detach().double().clamp(1e-12, 1 - 1e-12)
        lg = torch.log(fs / (1 - fs))
        print(variant, "loss", loss.item(), "norm", float(total), "logits", lg.min().item(), lg.max().item())
    np.savez_compressed(os.path.join(OUT, "attn_b3_1s.npz"), **out)


if __name__ == "__main__":
    main()
