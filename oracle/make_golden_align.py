"""Generate tests/golden/align_b4_1s.npz by running the UNMODIFIED reference sentence-level alignment models
(AudioTextAlignByWord / AudioTextAlignByPhrase, models/audio_text_model.py:843-976) with align.DotProduct
(models/align.py), every sim_pooling class (models/sim_pooling.py) and MaxMarginRankingLoss (losses.py:226-264) on a
seeded synthetic batch.  Build container only:   python oracle/make_golden_align.py
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
# the eg_configs use scaled=False (sentence_level/*/cnn8rnn_w2v*_dp_amean_tmean.yaml): the logits are 22x those of
# the scaled frame-level head, so the embedding is sharpened by 6 (not 100 / 300) to keep them within +-10
SHARPEN = 6.0
CASE = dict(batch=4, n_samples=32000, max_phrases=3, n_tokens=6, seed=5, data_seed=9)
POOLINGS = {
    "AudioMeanTextMean": ("mean", "mean"), "AudioMeanTextSum": ("mean", "sum"), "AudioMaxTextMean": ("max", "mean"),
    "AudioMaxTextMax": ("max", "max"), "AudioMaxTextSum": ("max", "sum"), "AudioMaxTextMeanSum": ("max", "meansum"),
    "AudioLinearSoftTextMean": ("linear_softmax", "mean"), "AudioLinearSoftTextSum": ("linear_softmax", "sum"),
    "AudioExpSoftTextMean": ("exp_softmax", "mean"), "AudioExpSoftTextSum": ("exp_softmax", "sum"),
}


def main():
    ns = ref_shim.import_reference()
    import models.audio_text_model as atm
    import models.align as align
    import models.sim_pooling as sp
    import losses
    torch.manual_seed(0)
    sd = O.synth_state_dict(seed=CASE["seed"], sharpen=SHARPEN, perturb_bn=True)
    batch = O.synth_align_batch(CASE["batch"], CASE["n_samples"], CASE["max_phrases"], CASE["n_tokens"],
                                seed=CASE["data_seed"])
    out = {"text": batch["text"].numpy(), "phrases": batch["phrases"].numpy(),
           "phrases_num": np.array(batch["phrases_num"]),
           "waveform_checksum": np.array([batch["waveform"].double().sum().item()])}

    def build(level, pooling, scaled=False):
        cls = atm.AudioTextAlignByWord if level == "word" else atm.AudioTextAlignByPhrase
        m = cls(ns.Cnn8Rnn(32000), ns.EmbeddingAgg(O.VOCAB, 512), align.DotProduct(l2norm=False, scaled=scaled),
                getattr(sp, pooling)(), 512)
        m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
        return m

    def inputs(level):
        d = {"specaug": False, "waveform": batch["waveform"], "waveform_len": batch["waveform_len"],
             "output_matrix": True}
        if level == "word":
            d.update({"text": batch["text"].long(), "text_len": batch["text_len"]})
        else:
            d.update({"phrases": batch["phrases"].long(), "phrases_len": batch["phrases_len"],
                      "phrases_num": batch["phrases_num"], "text_key": "phrases"})
        return d

    for level in ("word", "phrase"):
        for pooling in POOLINGS:
            m = build(level, pooling).eval()
            with torch.no_grad():
                o = m(inputs(level))
            out[f"eval_sim/{level}/{pooling}"] = o["sim"].numpy()
            if pooling == "AudioMeanTextMean":
                out[f"eval_sim_matrix/{level}"] = o["sim_matrix"].numpy()
        m = build(level, "AudioMeanTextMean", scaled=True).eval()
        with torch.no_grad():
            out[f"eval_sim_scaled/{level}"] = m(inputs(level))["sim"].numpy()

    loss_fn = losses.MaxMarginRankingLoss(margin=1, fix_norm=True, lamda1=1)
    for level, pooling in (("word", "AudioMeanTextMean"), ("phrase", "AudioMeanTextMean"),
                           ("phrase", "AudioLinearSoftTextSum"), ("word", "AudioMaxTextMeanSum"),
                           ("phrase", "AudioExpSoftTextMean")):
        tag = f"{level}/{pooling}"
        m = build(level, pooling).train()
        with dropout_identity():
            o = m(inputs(level))
            loss = loss_fn(o)
            loss.backward()
        total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in m.parameters() if p.grad is not None))
        out[f"train_loss/{tag}"] = np.array(loss.item())
        out[f"train_total_norm/{tag}"] = np.array(float(total))
        out[f"train_sim/{tag}"] = o["sim"].detach().numpy()
        for n, p in m.named_parameters():
            out[f"grad_norm/{tag}/{n}"] = np.array(p.grad.double().pow(2).sum().sqrt().item())
            out[f"grad_sub/{tag}/{n}"] = subsample(p.grad, 128)
        print(tag, "loss", loss.item(), "norm", float(total))
    # the loss alone on a fixed matrix, both normalisations and an asymmetric weight
    g = torch.Generator().manual_seed(3)
    x = torch.rand(6, 6, generator=g)
    out["mmr_x"] = x.numpy()
    for fix in (True, False):
        for lam in (1.0, 0.5):
            xx = x.clone().requires_grad_(True)
            l = losses.MaxMarginRankingLoss(margin=0.4, fix_norm=fix, lamda1=lam)({"sim": xx})
            l.backward()
            out[f"mmr_loss/{int(fix)}/{lam}"] = np.array(l.item())
            out[f"mmr_grad/{int(fix)}/{lam}"] = xx.grad.numpy()
    sm = torch.as_tensor(out["eval_sim_matrix/word"]).double()
    lg = torch.log(sm / (1 - sm).clamp_min(1e-12))
    print("word logits range", lg.min().item(), lg.max().item())
    np.savez_compressed(os.path.join(OUT, "align_b4_1s.npz"), **out)


if __name__ == "__main__":
    main()
