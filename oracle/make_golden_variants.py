"""Generate tests/golden/variants_b3_1s.npz by running the UNMODIFIED reference on the constructor options of
BiEncoder / MultiTextBiEncoder / EmbeddingAgg that the run_strong.py / run_weak_phrase.py config surface reaches:
  multi_proj ......... MultiTextBiEncoder(add_proj=True)                      models/audio_text_model.py:150-151,180-185
  multi_gating_proj .. MultiTextBiEncoder(add_proj=True, cross_encoder=CrossAttentionGating) + DotProduct("token")
                                                                              models/audio_text_model.py:166-185
  multi_upsample ..... MultiTextBiEncoder(upsample=True)                      models/audio_text_model.py:216-223
  bi_proj_upsample ... BiEncoder(add_proj=True, upsample=True)                models/audio_text_model.py:79-97
  bi_attnagg ......... BiEncoder with EmbeddingAgg(aggregation="attention")   models/text_encoder.py:46-58,84-85
Build container only:   python oracle/make_golden_variants.py      TEST INFRASTRUCTURE ONLY (see make_golden.py)."""
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
VARIANTS = ("multi_proj", "multi_gating_proj", "multi_upsample", "bi_proj_upsample", "bi_attnagg")


def build(ns, variant):
    import models.audio_text_model as atm
    import models.cross_encoder as ce
    import models.match as match
    import models.text_encoder as te
    enc = ns.Cnn8Rnn(32000)
    if variant == "multi_proj":
        m = atm.MultiTextBiEncoder(enc, te.EmbeddingAgg(O.VOCAB, 512), match.DotProduct(), 512,
                                   text_forward_keys=["text"], add_proj=True)
    elif variant == "multi_gating_proj":
        m = atm.MultiTextBiEncoder(enc, te.EmbeddingAgg(O.VOCAB, 512), match.DotProduct(text_level="token"), 512,
                                   text_forward_keys=["text"], add_proj=True,
                                   cross_encoder=ce.CrossAttentionGating(512))
    elif variant == "multi_upsample":
        m = atm.MultiTextBiEncoder(enc, te.EmbeddingAgg(O.VOCAB, 512), match.DotProduct(), 512,
                                   text_forward_keys=["text"], upsample=True)
    elif variant == "bi_proj_upsample":
        m = atm.BiEncoder(enc, te.EmbeddingAgg(O.VOCAB, 512), match.DotProduct(), 512, add_proj=True, upsample=True)
    else:
        m = atm.BiEncoder(enc, te.EmbeddingAgg(O.VOCAB, 512, aggregation="attention"), match.DotProduct(), 512)
    m.load_state_dict({k: v.clone() for k, v in O.variant_state(variant).items()}, strict=True)
    return m


def main():
    ns = ref_shim.import_reference()
    import losses
    torch.manual_seed(0)
    out = {}
    for variant in VARIANTS:
        batch = O.variant_batch(variant)
        multi = variant.startswith("multi")

        def inputs():
            d = {"specaug": False}
            d.update({k: (v.long() if k == "text" else v.float()) if isinstance(v, torch.Tensor) else v
                      for k, v in batch.items()})
            return d

        m = build(ns, variant).eval()
        with torch.no_grad():
            o = m(inputs())
            # the yardstick for the bf16 mode: the same reference forward under torch.autocast(bfloat16)
            with torch.autocast("cpu", dtype=torch.bfloat16):
                oa = m(inputs())
        out[f"eval_frame_sim_autocast/{variant}"] = oa["frame_sim"].float().numpy()
        print(variant, "reference autocast-bf16 vs fp32 frame_sim max abs",
              (oa["frame_sim"].float() - o["frame_sim"]).abs().max().item())
        out[f"text/{variant}"] = batch["text"].numpy()
        out[f"eval_frame_sim/{variant}"] = o["frame_sim"].numpy()
        out[f"eval_length/{variant}"] = np.asarray(o["length"])
        if multi:
            out[f"eval_clip_sim/{variant}"] = o["clip_sim"].numpy()
        m = build(ns, variant).train()
        with dropout_identity():
            o = m(inputs())
            if multi:
                o["label"] = batch["label"].float()
                loss = losses.ClipBceLoss()(o)
            else:
                T = min(o["frame_sim"].shape[1], batch["label"].shape[1])      # Runner.forward, run_strong.py:107-118
                o.update({"frame_sim": o["frame_sim"][..., :T], "label": batch["label"][..., :T].float(),
                          "length": torch.clamp(torch.as_tensor(o["length"]), 1, T)})
                loss = losses.FrameBceLoss()(o)
            loss.backward()
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
        out[f"train_loss/{variant}"] = np.array(loss.item())
        out[f"train_total_norm/{variant}"] = np.array(float(total))
        for n, p in m.named_parameters():
            if p.grad is None:
                continue
            out[f"grad_norm/{variant}/{n}"] = np.array(p.grad.double().pow(2).sum().sqrt().item())
            out[f"grad_sub/{variant}/{n}"] = subsample(p.grad, 128)
        fs = torch.as_tensor(out[f"eval_frame_sim/{variant}"]).double().clamp(1e-12, 1 - 1e-12)
        lg = torch.log(fs / (1 - fs))
        print(variant, "loss", loss.item(), "norm", float(total), "logits", lg.min().item(), lg.max().item(),
              "std", lg.std().item())
    np.savez_compressed(os.path.join(OUT, "variants_b3_1s.npz"), **out)


if __name__ == "__main__":
    main()
