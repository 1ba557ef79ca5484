"""Parity at BASELINE.json's FULL size (bs = 64, 10 s @ 32 kHz clips, 8-token phrases).

  * the fp32 train step (FusedTrainStep) at 10 s with B = 8 and B = 64 (ragged) against the CPU oracle's train step:
    loss, total norm, every parameter's gradient (cosine >= 0.999 on the FULL tensors) and the post-Adam parameters;
    the fp32 eval forward against the oracle on all 64 rows (1e-3, north_star)
  * the bf16 train step against the ORACLE with explicit floors, and against what the unmodified reference itself does
    under torch.autocast(bfloat16) on the same inputs (tests/golden/autocast_b64_10s.npz, oracle/make_golden_autocast.py)

and the size-independent properties kept from round 1:

  * eval forward is per-clip independent (BatchNorm uses running statistics): rows of the 64-clip batch equal the same
    clips run as a batch of 8, and clip 0 equals the CPU oracle on that clip alone (fp32: 1e-3, north_star)
  * bf16 tensor path vs fp32 path on the same weights: frame_sim within 1e-2 (north_star bf16 bar) for probabilities
    spanning (0.2, 0.8); the bf16 error is RELATIVE (0.66 % of the audio embedding after 8 conv layers and 250 GRU
    steps, measured with scripts/stage_err.py), so with the x300-sharpened weights of the fp32 checks (logits of
    +-10 and more) the same relative error shows as up to 3e-2 in probability — asserted as such, not hidden
  * full train step (fwd + bwd + clip + Adam), bf16 tcgen05 path vs fp32 CUDA-core path: loss, total gradient norm
    and per-parameter gradient direction agree; ragged lengths / padded tails included"""
import os

import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import GOLDEN, build_model, cosine, sub

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


def _oracle_train_step(sd, batch):
    sd_ref = {k: v.clone() for k, v in sd.items()}
    loss, grads, norm = O.train_step(sd_ref, batch, O.AdamState(O.trainable_keys()), dropout=False, fast_gru=True)
    return loss.item(), grads, norm.item(), sd_ref


def _full_cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten().to(a.device)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-300))


@pytest.mark.parametrize("nclips", [8, 64])
def test_full_length_fp32_train_step_matches_oracle(nclips):
    """The headline shape pinned to the oracle: 10 s clips, ragged lengths and token counts,
    B = 8 and the full B = 64, fp32 mode, dropout off: loss 2e-3, total norm 5e-3, every parameter gradient
    cosine >= 0.999 and norm within 1 % on the FULL tensors, post-Adam parameters."""
    from texttoaudiogrounding_b200.train import FusedTrainStep
    sd = O.synth_state_dict(seed=1, sharpen=30.0, perturb_bn=True)
    batch = {k: v[:nclips] for k, v in _batch().items()}
    ref_loss, ref_grads, ref_norm, sd_post = _oracle_train_step(sd, batch)
    model = build_model(sd, "fp32").train()
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
    loss = ts.step(batch).item()
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss, ref_loss, rtol=2e-3)
    np.testing.assert_allclose(ts.norm_out.item(), ref_norm, rtol=5e-3)
    named = dict(model.named_parameters())
    for k, rg in ref_grads.items():
        g = named[k].grad
        c = _full_cosine(g, rg)
        assert c >= 0.999, (k, c)
        gn, rn = float(g.double().norm()), float(rg.double().norm())
        assert abs(gn - rn) <= 1e-2 * rn + 1e-9, (k, gn, rn)
        # post-Adam parameters where the gradient is not noise-level (Adam's first step is lr * sign(g))
        solid = (rg.abs() > 1e-2 * rg.abs().max()).to(g.device)
        diff = (named[k].detach() - sd_post[k].to(g.device)).abs()[solid]
        assert float(diff.max()) <= 2e-4, (k, float(diff.max()))
    for k, v in model.named_buffers():
        if "running_" in k:
            np.testing.assert_allclose(v.cpu().numpy(), sd_post[k].numpy(), rtol=2e-3, atol=2e-4, err_msg=k)


def test_full_size_fp32_eval_matches_oracle_on_all_rows():
    sd = O.synth_state_dict(seed=3, sharpen=300.0, perturb_bn=True)
    batch = _batch()
    with torch.no_grad():
        ref = O.runner_forward({k: v.clone() for k, v in sd.items()}, batch, training=False, fast_gru=True)["frame_sim"]
    model = build_model(sd, "fp32").eval()
    with torch.no_grad():
        full = model(_to_dev(batch))["frame_sim"].cpu()
    assert full.shape == ref.shape == (B, 250)
    assert (full - ref).abs().max().item() <= 1e-3                       # north_star fp32 bar, all 64 x 250 frames
    assert ref.max().item() - ref.min().item() > 0.9                     # not vacuous
    # the committed fp32 run of the unmodified reference agrees with both
    g = np.load(os.path.join(GOLDEN, "autocast_b64_10s.npz"))
    assert np.abs(full.numpy() - g["eval_frame_sim_fp32"]).max() <= 1e-3
    assert np.abs(ref.numpy() - g["eval_frame_sim_fp32"]).max() <= 1e-4


def test_full_size_bf16_against_oracle_and_against_the_reference_under_autocast():
    """bf16 mode at the full size.  (1) eval, x300-sharpened weights (logits beyond +-10): no bf16 implementation can
    hold an absolute 1e-2 there — the unmodified reference under torch.autocast(bfloat16) is 1.7e-1 off its own
    fp32 run on these inputs (fixture); this repository's bf16 mode must be at least as close as that, and within 5e-2.
    (2) train step vs the ORACLE: loss 3e-2, total norm 5e-2, every parameter's gradient cosine >= 0.95 (bn0.weight: 0.93,
    see below) and not worse than what the reference's own autocast run achieves for that parameter minus 0.01 (0.02)."""
    from texttoaudiogrounding_b200.train import FusedTrainStep
    g = np.load(os.path.join(GOLDEN, "autocast_b64_10s.npz"))
    batch = _batch()
    sd = O.synth_state_dict(seed=3, sharpen=300.0, perturb_bn=True)
    model = build_model(sd, "bf16").eval()
    with torch.no_grad():
        full = model(_to_dev(batch))["frame_sim"].cpu().numpy()
    ours = np.abs(full - g["eval_frame_sim_fp32"]).max()
    theirs = np.abs(g["eval_frame_sim_autocast"] - g["eval_frame_sim_fp32"]).max()
    assert ours <= 5e-2 and ours <= theirs, (ours, theirs)
    del model
    sd = O.synth_state_dict(seed=1, sharpen=30.0, perturb_bn=True)
    ref_loss, ref_grads, ref_norm, _ = _oracle_train_step(sd, batch)
    np.testing.assert_allclose(ref_loss, g["train_loss_fp32"].item(), rtol=1e-4)      # oracle == reference (fp32)
    model = build_model(sd, "bf16").train()
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
    loss = ts.step(batch).item()
    torch.cuda.synchronize()
    assert abs(loss - ref_loss) <= 3e-2 * abs(ref_loss), (loss, ref_loss)
    assert abs(ts.norm_out.item() - ref_norm) <= 5e-2 * ref_norm, (ts.norm_out.item(), ref_norm)
    named = dict(model.named_parameters())
    report = {}
    for k, rg in ref_grads.items():
        c = _full_cosine(named[k].grad, rg)
        report[k] = (round(c, 4), round(float(g[f"grad_cosine/{k}"]), 4))
    print("bf16 gradient cosine vs oracle (ours, reference-under-autocast):",
          sorted(report.items(), key=lambda kv: kv[1][0])[:8])
    # profiles/r2_bn0_grad_probe.txt: over weight / data seeds the cosine of bn0.weight (and conv_block1.conv1.weight)
    # moves between 0.945 and 0.98 for ANY variant of the block-1 kernels — it is set by the bf16 rounding noise of the whole
    # backward chain (the reference's own autocast run sits at 0.9515 here), so its floor is 0.93 / reference - 0.02
    for k, (c, ref_c) in report.items():
        floor, slack = (0.93, 0.02) if k.endswith("bn0.weight") else (0.95, 0.01)
        assert c >= floor and c >= ref_c - slack, (k, c, ref_c)


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
