"""Constructor options of BiEncoder / MultiTextBiEncoder / EmbeddingAgg that the run_strong.py / run_weak_phrase.py
config surface reaches (reference models/audio_text_model.py:35-46, 79-97, 150-151, 166-185, 216-223;
models/text_encoder.py:46-58, 84-85): add_proj, cross_encoder inside MultiTextBiEncoder, upsample=True,
aggregation="attention".  CPU: the oracle restatement against the fixture generated from the unmodified reference
(oracle/make_golden_variants.py).  GPU: the mirrored modules (C-ABI kernels) against the same fixture.
Tolerance on the probability tensors: 1e-3 (fp32, north_star)."""
import os

import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import GOLDEN, cosine, sub

VARIANTS = ("multi_proj", "multi_gating_proj", "multi_upsample", "bi_proj_upsample", "bi_attnagg")


def load(variant):
    g = np.load(os.path.join(GOLDEN, "variants_b3_1s.npz"))
    sd = O.variant_state(variant)
    batch = O.variant_batch(variant)
    assert np.array_equal(batch["text"].numpy(), g[f"text/{variant}"])
    return g, sd, batch


def _oracle_loss(variant, out, batch):
    if variant.startswith("multi"):
        return O.clip_bce_loss(out["clip_sim"], batch["label"].float())
    T = min(out["frame_sim"].shape[1], batch["label"].shape[1])
    return O.frame_bce_loss({"frame_sim": out["frame_sim"][..., :T], "label": batch["label"][..., :T].float(),
                             "length": torch.clamp(torch.as_tensor(out["length"]), 1, T)})


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_variant_matches_reference(variant):
    g, sd, batch = load(variant)
    with torch.no_grad():
        out = O.variant_forward(variant, sd, batch, training=False)
    assert np.abs(out["frame_sim"].numpy() - g[f"eval_frame_sim/{variant}"]).max() <= 1e-4
    assert np.array_equal(np.asarray(out["length"]), g[f"eval_length/{variant}"])
    if variant.startswith("multi"):
        np.testing.assert_allclose(out["clip_sim"].numpy(), g[f"eval_clip_sim/{variant}"], atol=1e-4)
    keys = [k[len(f"grad_norm/{variant}/"):] for k in g.files if k.startswith(f"grad_norm/{variant}/")]
    params = []
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
        params.append(sd[k])
    out = O.variant_forward(variant, sd, batch, training=True, dropout=False)
    loss = _oracle_loss(variant, out, batch)
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{variant}"].item(), rtol=1e-4)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, gr in zip(keys, grads):
        ref = g[f"grad_norm/{variant}/{k}"].item()
        gn = 0.0 if gr is None else gr.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 5e-3 * ref + 1e-6, (k, gn, ref)


def test_oracle_linear_upsample_is_torch_interpolate():
    x = torch.randn(3, 17, generator=torch.Generator().manual_seed(0))
    ref = torch.nn.functional.interpolate(x.unsqueeze(1), 68, mode="linear", align_corners=False).squeeze(1)
    assert (O.linear_upsample(x, 68) - ref).abs().max() < 1e-6
    x3 = torch.randn(2, 25, 5, generator=torch.Generator().manual_seed(1))
    for size in (100, 20, 16, 25):        # up, down (the MultiTextBiEncoder quirk), identity
        ref3 = torch.nn.functional.interpolate(x3.transpose(1, 2), size, mode="linear",
                                               align_corners=False).transpose(1, 2)
        assert (O.linear_upsample(x3, size) - ref3).abs().max() < 1e-6, size


# ------------------------------------------------------------------------------------------- GPU
def _build(variant, sd, dtype="fp32"):
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder, MultiTextBiEncoder
    from texttoaudiogrounding_b200.models.cross_encoder import CrossAttentionGating
    from texttoaudiogrounding_b200.models.match import DotProduct
    from texttoaudiogrounding_b200.models.text_encoder import EmbeddingAgg
    enc = Cnn8Rnn(32000, compute_dtype=dtype)
    if variant == "multi_proj":
        m = MultiTextBiEncoder(enc, EmbeddingAgg(O.VOCAB, 512), DotProduct(), 512, text_forward_keys=["text"],
                               add_proj=True)
    elif variant == "multi_gating_proj":
        m = MultiTextBiEncoder(enc, EmbeddingAgg(O.VOCAB, 512), DotProduct(text_level="token"), 512,
                               text_forward_keys=["text"], add_proj=True, cross_encoder=CrossAttentionGating(512))
    elif variant == "multi_upsample":
        m = MultiTextBiEncoder(enc, EmbeddingAgg(O.VOCAB, 512), DotProduct(), 512, text_forward_keys=["text"],
                               upsample=True)
    elif variant == "bi_proj_upsample":
        m = BiEncoder(enc, EmbeddingAgg(O.VOCAB, 512), DotProduct(), 512, add_proj=True, upsample=True)
    else:
        m = BiEncoder(enc, EmbeddingAgg(O.VOCAB, 512, aggregation="attention"), DotProduct(), 512)
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m.cuda()


def _forward(model, batch):
    b = {k: ((v.long() if k == "text" else v.float()).cuda() if isinstance(v, torch.Tensor) else v)
         for k, v in batch.items()}
    d = {"specaug": False}
    d.update(b)
    return model(d), b


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_variant_eval_matches_reference_golden(variant):
    g, sd, batch = load(variant)
    model = _build(variant, sd).eval()
    with torch.no_grad():
        out, _ = _forward(model, batch)
    assert out["frame_sim"].shape == g[f"eval_frame_sim/{variant}"].shape
    assert np.abs(out["frame_sim"].cpu().numpy() - g[f"eval_frame_sim/{variant}"]).max() <= 1e-3      # fp32 bar
    assert np.array_equal(np.asarray(out["length"].cpu()), g[f"eval_length/{variant}"])
    if variant.startswith("multi"):
        np.testing.assert_allclose(out["clip_sim"].cpu().numpy(), g[f"eval_clip_sim/{variant}"], atol=1e-3)
    model = _build(variant, sd, "bf16").eval()
    with torch.no_grad():
        out, _ = _forward(model, batch)
    # bf16 bar: 1e-2 (north_star) — or, where the fixture's logits are too steep for ANY bf16 implementation
    # (multi_gating_proj: the unmodified reference under torch.autocast(bfloat16) is itself 3e-2 away from its fp32
    # output, recorded in the fixture by oracle/make_golden_variants.py), no worse than the reference's own bf16 run
    ref_bf16 = np.abs(g[f"eval_frame_sim_autocast/{variant}"] - g[f"eval_frame_sim/{variant}"]).max()
    assert np.abs(out["frame_sim"].cpu().numpy() - g[f"eval_frame_sim/{variant}"]).max() <= max(1e-2, ref_bf16)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_variant_loss_and_gradients_match_reference_golden(variant):
    from texttoaudiogrounding_b200.losses import ClipBceLoss, FrameBceLoss
    g, sd, batch = load(variant)
    model = _build(variant, sd).train()
    model.audio_encoder.dropout_enabled = False
    out, b = _forward(model, batch)
    if variant.startswith("multi"):
        out["label"] = b["label"]
        loss = ClipBceLoss()(out)
    else:
        T = min(out["frame_sim"].shape[1], b["label"].shape[1])
        out.update({"frame_sim": out["frame_sim"][..., :T], "label": b["label"][..., :T],
                    "length": torch.clamp(out["length"], 1, T)})
        loss = FrameBceLoss()(out)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{variant}"].item(), rtol=1e-3)
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters() if p.grad is not None)).item()
    np.testing.assert_allclose(total, g[f"train_total_norm/{variant}"].item(), rtol=1e-2)
    for n, p in model.named_parameters():
        key = f"grad_norm/{variant}/{n}"
        if key not in g.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        ref = g[key].item()
        assert p.grad is not None, f"{n} received no gradient"
        gn = p.grad.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref)
        if ref > 1e-6:
            assert cosine(sub(p.grad, 128), g[f"grad_sub/{variant}/{n}"]) > 0.999, n


@pytest.mark.gpu
def test_attention_pooling_rejects_short_batch_like_reference():
    """generate_length_mask(lens) is max(lens) wide: the reference's masked_fill fails when no phrase fills the
    padded width (models/text_encoder.py:54-55); the mirror raises the same RuntimeError."""
    from texttoaudiogrounding_b200.models.text_encoder import AttentionPooling
    pool = AttentionPooling(512).cuda()
    x = torch.randn(2, 6, 512, device="cuda")
    with pytest.raises(RuntimeError):
        pool(x, torch.tensor([3, 4]))
    assert pool(x, torch.tensor([6, 2])).shape == (2, 512)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,size", [((3, 17), 68), ((2, 9, 5), 36), ((1, 1, 3), 2), ((2, 250, 32), 1000),
                                        ((2, 250, 32), 128), ((3, 25, 4), 16)])
def test_upsample_linear_matches_torch(shape, size):
    from texttoaudiogrounding_b200.models import nn_ops
    F = torch.nn.functional
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    if x.dim() == 2:
        ref = F.interpolate(x.unsqueeze(1), size, mode="linear", align_corners=False).squeeze(1)
    else:
        ref = F.interpolate(x.transpose(1, 2), size, mode="linear", align_corners=False).transpose(1, 2)
    w = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    (ref * w).sum().backward()
    xc = x.detach().cuda().requires_grad_(True)
    out = nn_ops.upsample_linear(xc, size)
    (out * w.cuda()).sum().backward()
    assert (out.detach().cpu() - ref.detach()).abs().max().item() <= 1e-6
    assert (xc.grad.cpu() - x.grad).abs().max().item() <= 1e-5
