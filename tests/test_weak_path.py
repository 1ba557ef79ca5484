"""SURVEY.md §8f rank 1 — the weakly supervised multi-phrase path: MultiTextBiEncoder + *_with_lens pooling +
ClipBceLoss / ClipFrameBceLoss.  CPU: the oracle restatement against the fixture generated from the unmodified
reference (oracle/make_golden_weak.py).  GPU: the CUDA kernels (through the C ABI) and the mirrored modules
against the oracle and the same fixture.  Tolerance on the probability tensors: 1e-3 (fp32, north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import GOLDEN, cosine, rel_err, sub

CASE = dict(batch=3, n_samples=32000, n_phrases=5, n_tokens=6, seed=4, data_seed=7)
POOLINGS = ("linear_softmax", "max", "mean", "exp_softmax")


def load():
    g = np.load(os.path.join(GOLDEN, "weak_b3_1s.npz"))
    sd = O.synth_state_dict(seed=CASE["seed"], sharpen=100.0, perturb_bn=True)
    batch = O.synth_weak_batch(CASE["batch"], CASE["n_samples"], CASE["n_phrases"], CASE["n_tokens"],
                               seed=CASE["data_seed"])
    assert np.array_equal(batch["text"].numpy(), g["text"])
    return g, sd, batch


# ------------------------------------------------------------------------------------------- CPU
def test_oracle_multitext_forward_matches_reference():
    g, sd, batch = load()
    with torch.no_grad():
        for pooling in POOLINGS:
            out = O.multitext_forward(sd, batch, pooling=pooling, training=False)
            np.testing.assert_allclose(out["clip_sim"].numpy(), g[f"eval_clip_sim/{pooling}"], atol=1e-4)
    assert np.abs(out["frame_sim"].numpy() - g["eval_frame_sim"]).max() <= 1e-3
    assert np.array_equal(out["length"].numpy(), g["eval_length"])


@pytest.mark.parametrize("tag", ["clip", "clipframe"])
def test_oracle_weak_losses_and_gradients_match_reference(tag):
    g, sd, batch = load()
    keys = O.trainable_keys()
    params = []
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
        params.append(sd[k])
    out = O.multitext_forward(sd, batch, training=True, dropout=False)
    out.update({k: v for k, v in batch.items() if k != "text"})
    loss = O.clip_bce_loss(out["clip_sim"], out["label"]) if tag == "clip" else O.clip_frame_bce_loss(out, 0.3)
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{tag}"].item(), rtol=1e-4)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, gr in zip(keys, grads):
        ref = g[f"grad_norm/{tag}/{k}"].item()
        gn = gr.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 5e-3 * ref + 1e-7, (k, gn, ref)


# ------------------------------------------------------------------------------------------- GPU
def _gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,n", [(2, 50, 5), (3, 37, 33), (1, 250, 64)])
def test_multi_dot_sigmoid_fwd_bwd_matches_autograd(B, T, n):
    from texttoaudiogrounding_b200.models.match import DotProduct
    D = 512
    a = (torch.randn(B, T, D, generator=_gen(1)) * 0.5).requires_grad_(True)
    s = (torch.randn(B, n, D, generator=_gen(2)) * 0.5).requires_grad_(True)
    s.data[0, 0] *= 40.0               # drive some probabilities into the clamp (zero gradient there)
    ref = torch.sigmoid(torch.einsum("btd,bnd->btn", a, s) / D ** 0.5).clamp(1e-7, 1.0)
    w = torch.randn(B, T, n, generator=_gen(3))
    (ref * w).sum().backward()
    ac, sc = a.detach().cuda().requires_grad_(True), s.detach().cuda().requires_grad_(True)
    out = DotProduct().forward_multi(ac, sc)
    (out * w.cuda()).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 1e-5
    assert rel_err(ac.grad.cpu(), a.grad) < 1e-4
    assert rel_err(sc.grad.cpu(), s.grad) < 1e-4


@pytest.mark.gpu
@pytest.mark.parametrize("pooling", POOLINGS)
def test_pool_with_lens_fwd_bwd_matches_oracle(pooling):
    from texttoaudiogrounding_b200.models.utils import pool_with_lens
    B, T, n = 4, 23, 7
    f = (torch.rand(B, T, n, generator=_gen(5)) * 0.98 + 0.01).requires_grad_(True)
    lens = torch.tensor([23, 9, 1, 17])
    ref = O.pool_with_lens(f, lens, pooling)
    w = torch.randn(B, n, generator=_gen(6))
    (ref * w).sum().backward()
    fc = f.detach().cuda().requires_grad_(True)
    out = pool_with_lens(fc, lens, pooling)
    (out * w.cuda()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(fc.grad.cpu().numpy(), f.grad.numpy(), rtol=1e-4, atol=1e-6)
    # 2-D input (one phrase per clip) takes the same path
    out2 = pool_with_lens(fc.detach()[:, :, 0].contiguous(), lens, pooling)
    np.testing.assert_allclose(out2.cpu().numpy(), ref.detach().numpy()[:, 0], rtol=1e-5, atol=1e-6)


def _build_multi(sd, pooling="linear_softmax", dtype="fp32"):
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models.audio_text_model import MultiTextBiEncoder
    from texttoaudiogrounding_b200.models.match import DotProduct
    from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
    m = MultiTextBiEncoder(Cnn8Rnn(32000, compute_dtype=dtype), EmbeddingAgg(O.VOCAB, 512), DotProduct(), 512,
                           text_forward_keys=["text"], pooling=pooling)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m.cuda()


def _weak_forward(model, batch):
    """Runner.forward of run_weak_phrase.py:39-60."""
    b = {k: ((v.long() if k == "text" else v.float()).cuda() if isinstance(v, torch.Tensor) else v)
         for k, v in batch.items()}
    d = {"specaug": False}
    d.update(b)
    out = model(d)
    out.update({k: v for k, v in b.items() if k != "text"})
    return out


@pytest.mark.gpu
def test_multitext_biencoder_eval_matches_reference_golden():
    g, sd, batch = load()
    for pooling in POOLINGS:
        model = _build_multi(sd, pooling).eval()
        with torch.no_grad():
            out = _weak_forward(model, batch)
        np.testing.assert_allclose(out["clip_sim"].cpu().numpy(), g[f"eval_clip_sim/{pooling}"], atol=1e-3)
    assert np.abs(out["frame_sim"].cpu().numpy() - g["eval_frame_sim"]).max() <= 1e-3      # north_star fp32 bar
    assert np.array_equal(out["length"].cpu().numpy(), g["eval_length"])
    # bf16 compute mode: <= 1e-2 (north_star bf16 bar)
    model = _build_multi(sd, "linear_softmax", "bf16").eval()
    with torch.no_grad():
        out = _weak_forward(model, batch)
    assert np.abs(out["frame_sim"].cpu().numpy() - g["eval_frame_sim"]).max() <= 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["clip", "clipframe"])
def test_multitext_weak_losses_and_gradients_match_reference_golden(tag):
    from texttoaudiogrounding_b200.losses import ClipBceLoss, ClipFrameBceLoss
    g, sd, batch = load()
    model = _build_multi(sd).train()
    model.audio_encoder.dropout_enabled = False
    out = _weak_forward(model, batch)
    loss = (ClipBceLoss() if tag == "clip" else ClipFrameBceLoss(frame_weight=0.3))(out)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{tag}"].item(), rtol=1e-3)
    np.testing.assert_allclose(out["clip_sim"].detach().cpu().numpy(), g[f"train_clip_sim/{tag}"], atol=1e-3)
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters())).item()
    np.testing.assert_allclose(total, g[f"train_total_norm/{tag}"].item(), rtol=1e-2)
    for n, p in model.named_parameters():
        ref = g[f"grad_norm/{tag}/{n}"].item()
        gn = p.grad.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref)
        assert cosine(sub(p.grad, 128), g[f"grad_sub/{tag}/{n}"]) > 0.999, n


@pytest.mark.gpu
def test_weak_fused_train_step_matches_reference_golden():
    """Production path of the weak runners: flat buffers + fused clip/Adam (and CUDA-graph replay) for
    MultiTextBiEncoder + ClipBceLoss, against the reference's loss / gradient norms and against its own eager run."""
    from texttoaudiogrounding_b200.train import WeakFusedTrainStep
    g, sd, batch = load()
    model = _build_multi(sd).train()
    model.audio_encoder.dropout_enabled = False
    ts = WeakFusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
    loss = ts.step(batch).item()
    torch.cuda.synchronize()
    np.testing.assert_allclose(loss, g["train_loss/clip"].item(), rtol=1e-3)
    np.testing.assert_allclose(ts.norm_out.item(), g["train_total_norm/clip"].item(), rtol=1e-2)
    np.testing.assert_allclose(ts.clip.cpu().numpy(), g["train_clip_sim/clip"], atol=1e-3)
    for n, p in model.named_parameters():
        ref = g[f"grad_norm/clip/{n}"].item()
        gn = p.grad.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref)
        assert cosine(sub(p.grad, 128), g[f"grad_sub/clip/{n}"]) > 0.999, n
    # graph replay follows the eager trajectory
    sd2 = O.synth_state_dict(seed=1, sharpen=1.0, perturb_bn=True)
    losses = {}
    for use_graph in (False, True):
        m = _build_multi(sd2).train()
        m.audio_encoder.dropout_enabled = False
        t2 = WeakFusedTrainStep(m, use_graph=use_graph)
        losses[use_graph] = [t2.step(batch).item() for _ in range(5)]
    np.testing.assert_allclose(losses[True], losses[False], rtol=5e-3)
    assert losses[False][-1] < losses[False][0]


@pytest.mark.gpu
def test_weak_fused_train_step_clip_frame_loss_matches_reference_golden():
    """ClipFrameBceLoss(frame_weight=0.3) inside the fused weak step (eager and graph-replayed)."""
    from texttoaudiogrounding_b200.train import WeakFusedTrainStep
    g, sd, batch = load()
    for use_graph in (False, True):
        model = _build_multi(sd).train()
        model.audio_encoder.dropout_enabled = False
        ts = WeakFusedTrainStep(model, frame_weight=0.3, lr=0.0, use_graph=use_graph)
        for _ in range(3 if use_graph else 1):           # lr = 0: every step sees the same weights
            loss = ts.step(batch).item()
        torch.cuda.synchronize()
        np.testing.assert_allclose(loss, g["train_loss/clipframe"].item(), rtol=1e-3)
        np.testing.assert_allclose(ts.norm_out.item(), g["train_total_norm/clipframe"].item(), rtol=1e-2)
        for n, p in model.named_parameters():
            ref = g[f"grad_norm/clipframe/{n}"].item()
            gn = p.grad.double().pow(2).sum().sqrt().item()
            assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref, use_graph)
