"""Parity at BASELINE.json's FULL size (bs = 64, 10 s @ 32 kHz clips, 8-token phrases), where the CPU oracle is too slow
to run the whole batch: size-independent properties of the path plus one clip against the oracle.

  * eval forward is per-clip independent (BatchNorm uses running statistics): rows of the 64-clip batch equal the same
    clips run as a batch of 8, and clip 0 equals the CPU oracle on that clip alone (fp32: 1e-3, north_star)
  * bf16 tensor path vs fp32 path on the same weights: frame_sim within 1e-2 (north_star bf16 bar) for probabilities
    spanning (0.2, 0.8); the bf16 error is RELATIVE (0.66 % of the audio embedding after 8 conv layers and 250 GRU
    steps, measured with scripts/stage_err.py), so with the x300-sharpened weights of the fp32 checks (logits of
    +-10 and more) the same relative error shows as up to 3e-2 in probability — asserted as such, not hidden
  * full train step (fwd + bwd + clip + Adam), bf16 tcgen05 path vs fp32 CUDA-core path: loss, total gradient norm
    and per-parameter gradient direction agree; ragged lengths / padded tails included"""
import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import build_model, cosine, sub

pytestmark = pytest.mark.gpu
B, L = 64, 320000


def _batch():
    batch = O.synth_batch(B, L, seed=21, ragged=True)           # lengths L * {1, .9, .75, .5}, token counts 3..8
    return batch


def _to_dev(batch):
    d = {"specaug": False}
    for k, v in batch.items():
        d[k] = v.cuda() if isinstance(v, torch.Tensor) else v
    return d


def test_full_size_eval_is_per_clip_independent_and_matches_oracle_on_one_clip():
    sd = O.synth_state_dict(seed=3, sharpen=300.0, perturb_bn=True)
    batch = _batch()
    out = {}
    for prec in ("fp32", "bf16"):
        model = build_model(sd, prec).eval()
        with torch.no_grad():
            full = model(_to_dev(batch))["frame_sim"].cpu()
            part = model(_to_dev({k: v[:8] for k, v in batch.items()}))["frame_sim"].cpu()
        assert full.shape == (B, 250)
        tol = 1e-3 if prec == "fp32" else 1e-2
        assert (full[:8] - part).abs().max().item() <= tol, prec
        out[prec] = full
    assert (out["bf16"] - out["fp32"]).abs().max().item() <= 5e-2          # x300 stress weights, see the docstring
    one = {k: v[:1] for k, v in batch.items()}
    with torch.no_grad():
        ref = O.runner_forward({k: v.clone() for k, v in sd.items()}, one, training=False)["frame_sim"]
    assert (out["fp32"][:1, :ref.shape[1]] - ref).abs().max().item() <= 1e-3
    # the logits are not vacuous: the probabilities of this clip span most of (0, 1)
    assert ref.max().item() - ref.min().item() > 0.4
    # north_star bf16 bar at a realistic logit amplitude (x30: probabilities within about (0.2, 0.8))
    sd = O.synth_state_dict(seed=3, sharpen=30.0, perturb_bn=True)
    res = {}
    for prec in ("fp32", "bf16"):
        model = build_model(sd, prec).eval()
        with torch.no_grad():
            res[prec] = model(_to_dev(batch))["frame_sim"].cpu()
    assert res["fp32"].max().item() - res["fp32"].min().item() > 0.3
    assert (res["bf16"] - res["fp32"]).abs().max().item() <= 1e-2


def test_full_size_train_step_bf16_tensor_path_agrees_with_fp32_path():
    from texttoaudiogrounding_b200.train import FusedTrainStep
    sd = O.synth_state_dict(seed=1, sharpen=30.0, perturb_bn=True)
    batch = _batch()
    res = {}
    for prec in ("fp32", "bf16"):
        model = build_model(sd, prec).train()
        model.audio_encoder.dropout_enabled = False
        ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
        loss = ts.step(batch).item()
        torch.cuda.synchronize()
        res[prec] = (loss, ts.norm_out.item(), {n: sub(p.grad, 2048) for n, p in model.named_parameters()})
        del ts, model
        torch.cuda.empty_cache()
    (l32, n32, g32), (l16, n16, g16) = res["fp32"], res["bf16"]
    assert np.isfinite(l32) and np.isfinite(l16)
    assert abs(l16 - l32) <= 3e-2 * abs(l32), (l16, l32)
    assert abs(n16 - n32) <= 5e-2 * n32, (n16, n32)
    # measured (scripts/fullsize_grad_agreement.py): every parameter >= 0.9837 except bn0.weight at 0.954 — the gradient
    # of the input normalisation collects the bf16 rounding of all eight conv layers; run-to-run spread is 1e-13
    for n in g32:
        floor = 0.93 if ".bn0." in n else 0.97
        assert cosine(g16[n], g32[n]) > floor, (n, cosine(g16[n], g32[n]))
