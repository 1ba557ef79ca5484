"""SURVEY.md §8f rank 2 — attention-type heads (BASELINE.json configs[3]): text_encoder.SelfAttention,
match.CrossAttention, cross_encoder.CrossAttentionGating + DotProduct(text_level="token") inside BiEncoder.
CPU: the oracle restatement against the fixture generated from the unmodified reference
(oracle/make_golden_attn.py).  GPU: the CUDA kernels (through the C ABI) against torch autograd of the same op, and the
mirrored modules against the fixture.  Tolerance on frame_sim: 1e-3 (fp32, north_star) / 1e-2 (bf16 audio encoder)."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import tag_oracle as O
from helpers import GOLDEN, cosine, rel_err, sub

CASE = dict(batch=3, n_samples=32000, n_tokens=7, seed=6, data_seed=11, attn_seed=13)
VARIANTS = ("crossattn", "gating")


def load(variant):
    g = np.load(os.path.join(GOLDEN, "attn_b3_1s.npz"))
    sd = O.attn_case_state(variant, CASE["seed"], CASE["attn_seed"])
    batch = O.synth_batch(CASE["batch"], CASE["n_samples"], CASE["n_tokens"], seed=CASE["data_seed"], ragged=True)
    assert np.array_equal(batch["text"].numpy(), g["text"])
    return g, sd, batch


def trainable(sd):
    skip = ("running_", "num_batches", "pe.pe", "spectrogram.window", "mel_scale.fb")
    return [k for k in sd if not any(s in k for s in skip)]


# ------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_attn_forward_matches_reference(variant):
    g, sd, batch = load(variant)
    with torch.no_grad():
        out = O.attn_biencoder_forward(sd, batch, variant)
    np.testing.assert_allclose(out["seq_emb"].numpy(), g[f"eval_seq_emb/{variant}"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(out["token_emb"].numpy(), g[f"eval_token_emb/{variant}"], rtol=1e-4, atol=1e-4)
    assert np.abs(out["frame_sim"].numpy() - g[f"eval_frame_sim/{variant}"]).max() <= 1e-4
    assert np.array_equal(out["length"].numpy(), g[f"eval_length/{variant}"])


@pytest.mark.parametrize("variant", VARIANTS)
def test_oracle_attn_loss_and_gradients_match_reference(variant):
    g, sd, batch = load(variant)
    keys = trainable(sd)
    params = []
    for k in keys:
        sd[k] = sd[k].detach().requires_grad_(True)
        params.append(sd[k])
    out = O.attn_biencoder_forward(sd, batch, variant, training=True, dropout=False)
    T = out["frame_sim"].shape[1]
    loss = O.frame_bce_tensor(out["frame_sim"], batch["label"][:, :T], torch.as_tensor(out["length"]).clamp(1, T))
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{variant}"].item(), rtol=1e-4)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for k, gr in zip(keys, grads):
        ref = g[f"grad_norm/{variant}/{k}"].item()
        gn = gr.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 5e-3 * ref + 1e-6, (k, gn, ref)


# ------------------------------------------------------------------------------------------- GPU: kernels
def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _close(got, want, tol=2e-4):
    scale = max(want.abs().max().item(), 1e-6)
    assert (got.detach().cpu() - want.detach()).abs().max().item() <= tol * scale


@pytest.mark.gpu
@pytest.mark.parametrize("R,Cin,Cout", [(37, 512, 512), (576, 512, 1536), (9, 1024, 64),
                                        (2000, 512, 512), (1300, 1024, 192)])        # last two: tensor-core route
def test_linear_fwd_bwd_matches_autograd(R, Cin, Cout):
    from texttoaudiogrounding_b200.models import nn_ops
    x = torch.randn(R, Cin, generator=_gen(1)).requires_grad_(True)
    w = (torch.randn(Cout, Cin, generator=_gen(2)) * 0.05).requires_grad_(True)
    b = torch.randn(Cout, generator=_gen(3)).requires_grad_(True)
    dy = torch.randn(R, Cout, generator=_gen(4))
    F.linear(x, w, b).backward(dy)
    xc, wc, bc = (t.detach().cuda().requires_grad_(True) for t in (x, w, b))
    y = nn_ops.linear(xc, wc, bc)
    y.backward(dy.cuda())
    _close(y, F.linear(x, w, b))
    _close(xc.grad, x.grad)
    _close(wc.grad, w.grad)
    _close(bc.grad, b.grad)


@pytest.mark.gpu
@pytest.mark.parametrize("B,Lq,Lk,heads", [(3, 8, 8, 8), (2, 50, 7, 8), (2, 5, 40, 4), (1, 3, 9, 16)])
def test_multi_head_attention_matches_torch(B, Lq, Lk, heads):
    from texttoaudiogrounding_b200.models import nn_ops
    E = 512
    torch.manual_seed(5)
    mha = torch.nn.MultiheadAttention(E, heads, 0.0, batch_first=True)
    with torch.no_grad():
        mha.in_proj_bias.uniform_(-0.2, 0.2)
        mha.out_proj.bias.uniform_(-0.2, 0.2)
    self_attn = Lq == Lk
    xq = torch.randn(B, Lq, E, generator=_gen(1)).requires_grad_(True)
    xk = xq if self_attn else torch.randn(B, Lk, E, generator=_gen(2)).requires_grad_(True)
    klen = torch.randint(1, Lk + 1, (B,), generator=_gen(3))
    klen[0] = Lk
    mask = ~O.generate_length_mask(klen, Lk)
    ref, _ = mha(xq, xk, xk, key_padding_mask=mask)
    dy = torch.randn(B, Lq, E, generator=_gen(4))
    ref.backward(dy)
    ref_grads = {n: p.grad.clone() for n, p in mha.named_parameters()}
    gxq, gxk = xq.grad.clone(), xk.grad.clone()
    mha.zero_grad()
    mc = mha.cuda()
    cq = xq.detach().cuda().requires_grad_(True)
    ck = cq if self_attn else xk.detach().cuda().requires_grad_(True)
    out = nn_ops.multi_head_attention(mc, cq, ck, ck, klen.cuda(), training=False)
    out.backward(dy.cuda())
    _close(out, ref)
    _close(cq.grad, gxq)
    if not self_attn:
        _close(ck.grad, gxk)
    for n, p in mc.named_parameters():
        _close(p.grad, ref_grads[n])


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,N", [(3, 50, 7), (2, 37, 1), (2, 70, 24), (5, 250, 7)])
def test_cross_attention_gating_matches_oracle(B, T, N):
    from texttoaudiogrounding_b200.models.cross_encoder import CrossAttentionGating
    E = 512
    sd = {k[len("cross_encoder."):]: v for k, v in O.synth_attn_state(3, ("gating",), gain=2.0).items()}
    a = (torch.randn(B, T, E, generator=_gen(1)) * 0.7).requires_grad_(True)
    x = torch.randn(B, N, E, generator=_gen(2)).requires_grad_(True)
    alen = torch.randint(1, T + 1, (B,), generator=_gen(3))
    alen[0] = T
    tlen = torch.randint(1, N + 1, (B,), generator=_gen(4))
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    u, s = O.cross_attention_gating({"cross_encoder." + k: v for k, v in params.items()}, a, x, alen, tlen)
    wu, ws = torch.randn(B, T, E, generator=_gen(5)), torch.randn(B, T, E, generator=_gen(6))
    ((u * wu).sum() + (s * ws).sum()).backward()
    m = CrossAttentionGating(E)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    ac, xc = a.detach().cuda().requires_grad_(True), x.detach().cuda().requires_grad_(True)
    out = m({"audio_emb": ac, "text_emb": {"token_emb": xc}, "audio_len": alen, "text_len": tlen})
    ((out["audio_emb"] * wu.cuda()).sum() + (out["text_emb"]["token_emb"] * ws.cuda()).sum()).backward()
    _close(out["audio_emb"], u)
    _close(out["text_emb"]["token_emb"], s)
    _close(ac.grad, a.grad)
    _close(xc.grad, x.grad)
    for n, p in m.named_parameters():
        _close(p.grad, params[n].grad, 5e-4)


@pytest.mark.gpu
def test_cross_attention_match_and_token_dot_match_oracle():
    from texttoaudiogrounding_b200.models.match import CrossAttention, DotProduct
    B, T, N, E = 3, 41, 6, 512
    sd = {k[len("match_fn."):]: v for k, v in O.synth_attn_state(4, ("crossattn",)).items()}
    a = torch.randn(B, T, E, generator=_gen(1)).requires_grad_(True)
    x = torch.randn(B, N, E, generator=_gen(2)).requires_grad_(True)
    tlen = torch.tensor([N, 2, 4])
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = O.cross_attention_match({"match_fn." + k: v for k, v in params.items()}, a, x, tlen)
    w = torch.randn(B, T, generator=_gen(3))
    (ref * w).sum().backward()
    m = CrossAttention(E, O.HEADS, 0.0)
    m.load_state_dict(sd, strict=True)
    m = m.cuda()
    ac, xc = a.detach().cuda().requires_grad_(True), x.detach().cuda().requires_grad_(True)
    out = m({"audio_emb": ac, "text_emb": {"token_emb": xc}, "text_len": tlen})
    (out * w.cuda()).sum().backward()
    _close(out, ref, 1e-5)
    _close(ac.grad, a.grad)
    _close(xc.grad, x.grad)
    for n, p in m.named_parameters():
        _close(p.grad, params[n].grad, 5e-4)
    # token-level DotProduct on per-frame text
    t = (torch.randn(B, T, E, generator=_gen(7)) * 0.3).requires_grad_(True)
    a2 = a.detach().clone().requires_grad_(True)
    ref2 = torch.sigmoid((a2 * t).sum(-1) / math.sqrt(E)).clamp(1e-7, 1.0)
    (ref2 * w).sum().backward()
    ac2, tc = a2.detach().cuda().requires_grad_(True), t.detach().cuda().requires_grad_(True)
    out2 = DotProduct(text_level="token")({"audio_emb": ac2, "text_emb": {"token_emb": tc}})
    (out2 * w.cuda()).sum().backward()
    _close(out2, ref2, 1e-5)
    _close(ac2.grad, a2.grad)
    _close(tc.grad, t.grad)


# ------------------------------------------------------------------------------------------- GPU: models
def _build(sd, variant, dtype="fp32"):
    from texttoaudiogrounding_b200.models.audio_encoder import Cnn8Rnn
    from texttoaudiogrounding_b200.models.audio_text_model import BiEncoder
    from texttoaudiogrounding_b200.models.cross_encoder import CrossAttentionGating
    from texttoaudiogrounding_b200.models.match import CrossAttention, DotProduct
    from texttoaudiogrounding_b200.models.text_encoder import SelfAttention
    text_encoder = SelfAttention(O.VOCAB, 512, O.HEADS, dropout=0.0)
    if variant == "crossattn":
        m = BiEncoder(Cnn8Rnn(32000, compute_dtype=dtype), text_encoder, CrossAttention(512, O.HEADS, 0.0), 512)
    else:
        m = BiEncoder(Cnn8Rnn(32000, compute_dtype=dtype), text_encoder, DotProduct(text_level="token"), 512,
                      cross_encoder=CrossAttentionGating(512))
    m.load_state_dict({k: v.clone() for k, v in sd.items()}, strict=True)
    return m.cuda()


def _inputs(batch):
    return {"specaug": False, "waveform": batch["waveform"].cuda(), "waveform_len": batch["waveform_len"],
            "text": batch["text"].long().cuda(), "text_len": batch["text_len"]}


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_attn_models_eval_match_reference_golden(variant):
    g, sd, batch = load(variant)
    model = _build(sd, variant).eval()
    with torch.no_grad():
        out = model(_inputs(batch))
        t = model.text_encoder(_inputs(batch))
    assert rel_err(t["seq_emb"].cpu(), g[f"eval_seq_emb/{variant}"]) < 1e-4
    assert rel_err(t["token_emb"].cpu(), g[f"eval_token_emb/{variant}"]) < 1e-4
    assert np.abs(out["frame_sim"].cpu().numpy() - g[f"eval_frame_sim/{variant}"]).max() <= 1e-3     # fp32 bar
    assert np.array_equal(np.asarray(out["length"].cpu()), g[f"eval_length/{variant}"])
    model = _build(sd, variant, "bf16").eval()
    with torch.no_grad():
        out = model(_inputs(batch))
    assert np.abs(out["frame_sim"].cpu().numpy() - g[f"eval_frame_sim/{variant}"]).max() <= 1e-2     # bf16 bar


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_attn_models_loss_and_gradients_match_reference_golden(variant):
    from texttoaudiogrounding_b200.losses import FrameBceLoss
    g, sd, batch = load(variant)
    model = _build(sd, variant).train()
    model.audio_encoder.dropout_enabled = False
    out = model(_inputs(batch))
    T = out["frame_sim"].shape[1]
    out["label"] = batch["label"][:, :T].cuda()
    out["length"] = torch.as_tensor(out["length"]).clamp(1, T)
    loss = FrameBceLoss()(out)
    loss.backward()
    np.testing.assert_allclose(loss.item(), g[f"train_loss/{variant}"].item(), rtol=1e-3)
    np.testing.assert_allclose(out["frame_sim"].detach().cpu().numpy(), g[f"train_frame_sim/{variant}"], atol=1e-3)
    total = torch.sqrt(sum((p.grad.double() ** 2).sum() for p in model.parameters())).item()
    np.testing.assert_allclose(total, g[f"train_total_norm/{variant}"].item(), rtol=1e-2)
    for n, p in model.named_parameters():
        ref = g[f"grad_norm/{variant}/{n}"].item()
        gn = p.grad.double().pow(2).sum().sqrt().item()
        assert abs(gn - ref) <= 1e-2 * ref + 1e-6, (n, gn, ref)
        assert cosine(sub(p.grad, 128), g[f"grad_sub/{variant}/{n}"]) > 0.999, n


@pytest.mark.gpu
@pytest.mark.parametrize("variant", VARIANTS)
def test_autograd_train_step_matches_torch_adam_and_replays_as_graph(variant):
    """AutogradTrainStep (flat buffers + fused clip/Adam, BASELINE.json configs[3] train loop): three steps eager ==
    three steps of train_step with torch.optim.Adam on an identical model; a graph-replaying instance (first step
    eager on the capture stream, second step captured, third replayed) follows the same trajectory."""
    from texttoaudiogrounding_b200.train import AutogradTrainStep, train_step
    g, sd, batch = load(variant)
    dev_batch = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}

    def fresh():
        m = _build(sd, variant).train()
        m.audio_encoder.dropout_enabled = False
        return m

    ref = fresh()
    opt = torch.optim.Adam(ref.parameters(), lr=1e-3)
    ref_losses = [train_step(ref, dev_batch, opt)[0].item() for _ in range(3)]
    np.testing.assert_allclose(ref_losses[0], g[f"train_loss/{variant}"].item(), rtol=1e-3)
    runs = {}
    for use_graph in (False, True):
        m = fresh()
        ts = AutogradTrainStep(m, lr=1e-3, max_grad_norm=1.0, use_graph=use_graph)
        runs[use_graph] = ([ts.step(batch).item() for _ in range(3)], m)
        if use_graph:
            assert len(ts._graphs) == 1
    p0 = {n: p.detach().clone() for n, p in fresh().named_parameters()}
    for use_graph, (losses, m) in runs.items():
        np.testing.assert_allclose(losses, ref_losses, rtol=2e-3, err_msg=f"graph={use_graph}")
        # Adam divides by sqrt(v): an element whose gradient is at the rounding-noise level (split-K atomics) may step
        # the other way, so single elements differ by up to 2 * lr * steps; the UPDATE vectors must agree as a whole
        du = torch.cat([(p.detach() - p0[n]).flatten() for n, p in m.named_parameters()]).double()
        dr = torch.cat([(q.detach() - p0[n]).flatten() for n, q in ref.named_parameters()]).double()
        assert (du - dr).abs().max().item() <= 2 * 1e-3 * 3 + 1e-6
        assert float((du * dr).sum() / (du.norm() * dr.norm())) > 0.99, use_graph
        assert float((du - dr).abs().mean() / dr.abs().mean()) < 0.05, use_graph
