"""SURVEY.md §8f rank 3 — sentence-level alignment: align.DotProduct + sim_pooling.* + AudioTextAlignBy{Word,Phrase}
+ MaxMarginRankingLoss.  CPU: the oracle restatement against the fixture generated from the unmodified reference
(oracle/make_golden_align.py).  GPU: the fused CUDA kernels (through the C ABI) and the mirrored modules against the
oracle and the same fixture.  Tolerance on probabilities / pooled similarities: 1e-3 (fp32, north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import GOLDEN, cosine, sub

CASE = dict(batch=4, n_samples=32000, max_phrases=3, n_tokens=6, seed=5, data_seed=9)
POOLINGS = {
    "AudioMeanTextMean": ("mean", "mean"), "AudioMeanTextSum": ("mean", "sum"), "AudioMaxTextMean": ("max", "mean"),
    "AudioMaxTextMax": ("max", "max"), "AudioMaxTextSum": ("max", "sum"), "AudioMaxTextMeanSum": ("max", "meansum"),
    "AudioLinearSoftTextMean": ("linear_softmax", "mean"), "AudioLinearSoftTextSum": ("linear_softmax", "sum"),
    "AudioExpSoftTextMean": ("exp_softmax", "mean"), "AudioExpSoftTextSum": ("exp_softmax", "sum"),
}
TRAIN_TAGS = [("word", "AudioMeanTextMean"), ("phrase", "AudioMeanTextMean"), ("phrase", "AudioLinearSoftTextSum"),
              ("word", "AudioMaxTextMeanSum"), ("phrase", "AudioExpSoftTextMean")]


def load():
    g = np.load(os.path.join(GOLDEN, "align_b4_1s.npz"))
    sd = O.synth_state_dict(seed=CASE["seed"], sharpen=6.0, perturb_bn=True)
    batch = O.synth_align_batch(CASE["batch"], CASE["n_samples"], CASE["max_phrases"], CASE["n_tokens"],
                                seed=CASE["data_seed"])
    assert np.array_equal(batch["phrases"].numpy(), g["phrases"])
    assert np.array_equal(np.array(batch["phrases_num"]), g["phrases_num"])
    return g, sd, batch


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("level", ["word", "phrase"])
def test_oracle_align_forward_matches_reference(level):
    g, sd, batch = load()
    with torch.no_grad():
        for name, (ap, tp) in POOLINGS.items():
            out = O.align_forward(sd, batch, level, ap, tp)
            np.testing.assert_allclose(out["sim"].numpy(), g[f"eval_sim/{level}/{name}"], rtol=1e-4, atol=1e-5)
        assert np.abs(out["sim_matrix"].numpy() - g[f"eval_sim_matrix/{level}"]).max() <= 1e-4
        out = O.align_forward(sd, batch, level, "mean", "mean", scaled=True)
        np.testing.assert_allclose(out["sim"].numpy(), g[f"eval_sim_scaled/{level}"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("level,pooling", TRAIN_TAGS[:3])
def test_oracle_align_loss_and_gradients_match_reference(level, pooling):
    g, sd, batch = load()
    tag = f"{level}/{pooling}"
    keys = O.trainable_keys()
    params = []
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
        params.append(sd[k])
    out = O.align_forward(sd, batch, level, *POOLINGS[pooling], training=True, dropout=False)
    loss = O.max_margin_ranking_loss(out["sim"])
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{tag}"].item(), rtol=1e-4)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, gr in zip(keys, grads):
        ref = g[f"grad_norm/{tag}/{k}"].item()
        gn = gr.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 5e-3 * ref + 1e-7, (k, gn, ref)


def test_oracle_max_margin_ranking_loss_matches_reference():
    g = np.load(os.path.join(GOLDEN, "align_b4_1s.npz"))
    for fix in (True, False):
        for lam in (1.0, 0.5):
            x = torch.as_tensor(g["mmr_x"]).requires_grad_(True)
            l = O.max_margin_ranking_loss(x, 0.4, fix, lam)
            l.backward()
            np.testing.assert_allclose(l.item(), g[f"mmr_loss/{int(fix)}/{lam}"].item(), rtol=1e-6)
            np.testing.assert_allclose(x.grad.numpy(), g[f"mmr_grad/{int(fix)}/{lam}"], atol=1e-7)


# ------------------------------------------------------------------------------------------- GPU
def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(POOLINGS))
@pytest.mark.parametrize("B,T,N,scaled", [(3, 37, 5, True), (4, 50, 6, False), (2, 250, 8, True), (9, 13, 9, True),
                                          (8, 250, 8, False), (12, 101, 11, True)])     # last two: tensor-core route
def test_fused_align_pool_fwd_bwd_matches_oracle(name, B, T, N, scaled):
    from texttoaudiogrounding_b200.models.align import DotProduct
    import texttoaudiogrounding_b200.models.sim_pooling as sp
    D = 512
    amp = 0.5 if scaled else 0.1
    a = (torch.randn(B, T, D, generator=_gen(1)) * amp).requires_grad_(True)
    x = (torch.randn(B, N, D, generator=_gen(2)) * amp).requires_grad_(True)
    x.data[0, 0] *= 60.0               # drive some probabilities into the clamp (zero gradient there)
    alen = torch.randint(1, T + 1, (B,), generator=_gen(3))
    alen[0] = T                        # the reference's max_with_lens needs max(len) == T
    tlen = torch.randint(1, N + 1, (B,), generator=_gen(4))
    ap, tp = POOLINGS[name]
    ref = O.sim_pooling(O.align_dot_product(a, x, scaled), alen, tlen, ap, tp)
    w = torch.randn(B, B, generator=_gen(5))
    (ref * w).sum().backward()
    ac, xc = a.detach().cuda().requires_grad_(True), x.detach().cuda().requires_grad_(True)
    sim = DotProduct(scaled=scaled)(ac, xc)
    assert tuple(sim.size()) == (B, B, T, N)
    out = getattr(sp, name)()({"sim": sim, "audio_len": alen, "text_len": tlen})
    (out * w.cuda()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=2e-5, atol=2e-6)
    for got, want in ((ac.grad, a.grad), (xc.grad, x.grad)):
        scale = want.abs().max().item()
        assert (got.cpu() - want).abs().max().item() <= 2e-4 * scale + 1e-7
    m = sim.materialize().cpu()
    # >= 1024 rows run the split-bf16 tensor-core GEMM: logits to ~2e-5 relative (|logit| reaches 15 here)
    assert (m - O.align_dot_product(a, x, scaled).detach()).abs().max().item() <= (1e-4 if B * T >= 1024 else 1e-5)


@pytest.mark.gpu
def test_max_margin_ranking_loss_matches_reference_golden():
    from texttoaudiogrounding_b200.losses import MaxMarginRankingLoss
    g = np.load(os.path.join(GOLDEN, "align_b4_1s.npz"))
    for fix in (True, False):
        for lam in (1.0, 0.5):
            x = torch.as_tensor(g["mmr_x"]).cuda().requires_grad_(True)
            l = MaxMarginRankingLoss(margin=0.4, fix_norm=fix, lamda1=lam)({"sim": x})
            (3.0 * l).backward()
            np.testing.assert_allclose(l.item(), g[f"mmr_loss/{int(fix)}/{lam}"].item(), rtol=1e-5)
            np.testing.assert_allclose(x.grad.cpu().numpy() / 3.0, g[f"mmr_grad/{int(fix)}/{lam}"], atol=1e-6)


def _build(sd, level, pooling, scaled=False, dtype="fp32"):
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models import audio_text_model as atm
    from texttoaudiogrounding_b200.models.align import DotProduct
    import texttoaudiogrounding_b200.models.sim_pooling as sp
    from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
    cls = atm.AudioTextAlignByWord if level == "word" else atm.AudioTextAlignByPhrase
    m = cls(Cnn8Rnn(32000, compute_dtype=dtype), EmbeddingAgg(O.VOCAB, 512), DotProduct(l2norm=False, scaled=scaled),
            getattr(sp, pooling)(), 512)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m.cuda()


def _inputs(batch, level):
    d = {"specaug": False, "waveform": batch["waveform"].cuda(), "waveform_len": batch["waveform_len"],
         "output_matrix": True}
    if level == "word":
        d.update({"text": batch["text"].long().cuda(), "text_len": batch["text_len"]})
    else:
        d.update({"phrases": batch["phrases"].long().cuda(), "phrases_len": batch["phrases_len"],
                  "phrases_num": batch["phrases_num"], "text_key": "phrases"})
    return d


@pytest.mark.gpu
@pytest.mark.parametrize("level", ["word", "phrase"])
def test_align_models_eval_match_reference_golden(level):
    g, sd, batch = load()
    for name in POOLINGS:
        model = _build(sd, level, name).eval()
        with torch.no_grad():
            out = model(_inputs(batch, level))
        ref = g[f"eval_sim/{level}/{name}"]
        assert np.abs(out["sim"].cpu().numpy() - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max()), name
    assert np.abs(out["sim_matrix"].cpu().numpy() - g[f"eval_sim_matrix/{level}"]).max() <= 1e-3   # north_star fp32
    model = _build(sd, level, "AudioMeanTextMean", scaled=True).eval()
    with torch.no_grad():
        out = model(_inputs(batch, level))
    assert np.abs(out["sim"].cpu().numpy() - g[f"eval_sim_scaled/{level}"]).max() <= 1e-3
    model = _build(sd, level, "AudioMeanTextMean", dtype="bf16").eval()                             # bf16 bar 1e-2
    with torch.no_grad():
        out = model(_inputs(batch, level))
    assert np.abs(out["sim"].cpu().numpy() - g[f"eval_sim/{level}/AudioMeanTextMean"]).max() <= 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("level,pooling", TRAIN_TAGS)
def test_align_models_loss_and_gradients_match_reference_golden(level, pooling):
    from texttoaudiogrounding_b200.losses import MaxMarginRankingLoss
    g, sd, batch = load()
    tag = f"{level}/{pooling}"
    model = _build(sd, level, pooling).train()
    model.audio_encoder.dropout_enabled = False
    out = model(_inputs(batch, level))
    loss = MaxMarginRankingLoss(margin=1, fix_norm=True, lamda1=1)(out)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{tag}"].item(), rtol=1e-3)
    np.testing.assert_allclose(out["sim"].detach().cpu().numpy(), g[f"train_sim/{tag}"], atol=1e-3)
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters())).item()
    np.testing.assert_allclose(total, g[f"train_total_norm/{tag}"].item(), rtol=1e-2)
    for n, p in model.named_parameters():
        ref = g[f"grad_norm/{tag}/{n}"].item()
        gn = p.grad.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref)
        assert cosine(sub(p.grad, 128), g[f"grad_sub/{tag}/{n}"]) > 0.999, n


@pytest.mark.gpu
@pytest.mark.parametrize("level,pooling", TRAIN_TAGS[:3])
def test_align_fused_train_step_matches_reference_golden(level, pooling):
    """Production path of the sentence-level runners: flat buffers + fused clip/Adam (+ CUDA-graph replay) for
    AudioTextAlignBy{Word,Phrase} + MaxMarginRankingLoss, against the reference's loss and gradient norms."""
    from texttoaudiogrounding_b200.train import AlignFusedTrainStep
    g, sd, batch = load()
    tag = f"{level}/{pooling}"
    for use_graph in (False, True):
        model = _build(sd, level, pooling).train()
        model.audio_encoder.dropout_enabled = False
        ts = AlignFusedTrainStep(model, margin=1, fix_norm=True, lamda1=1, lr=0.0, use_graph=use_graph)
        for _ in range(3 if use_graph else 1):           # lr = 0: every step sees the same weights
            loss = ts.step(batch).item()
        torch.cuda.synchronize()
        np.testing.assert_allclose(loss, g[f"train_loss/{tag}"].item(), rtol=1e-3)
        np.testing.assert_allclose(ts.sim.cpu().numpy(), g[f"train_sim/{tag}"], atol=1e-3)
        np.testing.assert_allclose(ts.norm_out.item(), g[f"train_total_norm/{tag}"].item(), rtol=1e-2)
        for n, p in model.named_parameters():
            ref = g[f"grad_norm/{tag}/{n}"].item()
            gn = p.grad.double().pow(2).sum().sqrt().item()
            assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref, use_graph)
