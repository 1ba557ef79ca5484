"""GPU parity tests: the CUDA path (through the C ABI) against the golden fixtures generated
from the unmodified reference and against the CPU oracle on the same seeded inputs.

Tolerances: fp32 mode  <= 1e-3 on frame_sim (north_star), stage tensors rel-err <= 2e-3;
            bf16 mode  <= 1e-2 on frame_sim (north_star)."""
import numpy as np
import pytest
import torch

from oracle import tag_oracle as O
from helpers import CASES, build_model, cosine, load_case, nchw, rel_err, sub

pytestmark = pytest.mark.gpu


def _eval_forward(model, batch):
    from texttoaudiogrounding_b200.train import runner_forward
    stages = {}
    model.audio_encoder._stages = stages
    model.eval()
    with torch.no_grad():
        out = runner_forward(model, dict(batch), "cuda", training=False)
    model.audio_encoder._stages = None
    torch.cuda.synchronize()
    return out, stages


@pytest.mark.parametrize("name", list(CASES))
def test_eval_forward_fp32_matches_reference_golden(name):
    g, sd, batch = load_case(name)
    model = build_model(sd, "fp32")
    out, st = _eval_forward(model, batch)
    db = st["logmel_db"].cpu().numpy()                       # [B,T0,64]
    ref_db = np.transpose(g["eval_logmel_db"], (0, 2, 1))    # reference: [B,64,T0]
    assert np.abs(db - ref_db).max() <= 5e-3, np.abs(db - ref_db).max()
    bn0 = st["bn0"].float().permute(0, 2, 1).unsqueeze(-1).contiguous()   # [B,64,T0,1]
    np.testing.assert_allclose(sub(bn0), g["eval_bn0_sub"], rtol=2e-3, atol=2e-3)
    for i in range(1, 5):
        got = sub(nchw(st[f"conv_block{i}"].float()))
        np.testing.assert_allclose(got, g[f"eval_conv_block{i}_sub"], rtol=2e-3, atol=2e-3,
                                   err_msg=f"conv_block{i}")
    B = batch["waveform"].shape[0]
    fc1 = st["fc1"].float().reshape(B, -1, 512)
    np.testing.assert_allclose(sub(fc1), g["eval_fc1_sub"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(st["rnn"].cpu().numpy(), g["eval_embedding"], atol=1e-3)
    fs = out["frame_sim"].cpu().numpy()
    assert np.abs(fs - g["eval_frame_sim"]).max() <= 1e-3, np.abs(fs - g["eval_frame_sim"]).max()
    logits = np.log(fs.astype(np.float64) / (1 - fs.astype(np.float64)))
    mask = np.abs(g["eval_logits"]) < 10          # away from sigmoid saturation
    np.testing.assert_allclose(logits[mask], g["eval_logits"][mask], atol=5e-3, rtol=2e-3)
    assert np.array_equal(out["length"].cpu().numpy(), g["eval_length"])


@pytest.mark.parametrize("name", list(CASES))
def test_eval_forward_bf16_within_contract(name):
    g, sd, batch = load_case(name)
    model = build_model(sd, "bf16")
    out, st = _eval_forward(model, batch)
    fs = out["frame_sim"].cpu().numpy()
    # north_star: <= 1e-2 on the frame-probability tensor for bf16
    assert np.abs(fs - g["eval_frame_sim"]).max() <= 1e-2, np.abs(fs - g["eval_frame_sim"]).max()
    assert rel_err(st["rnn"].cpu(), g["eval_embedding"]) < 3e-2


def _check_train_against_golden(g, grads, loss, total_norm, post, bufs, grad_rel, cos_min):
    np.testing.assert_allclose(loss, g["train_loss"].item(), rtol=grad_rel)
    np.testing.assert_allclose(total_norm, g["train_total_norm"].item(), rtol=max(grad_rel, 2e-3))
    for k in g["param_names"].tolist():
        gn = float(grads[k].double().pow(2).sum().sqrt())
        ref = g[f"grad_norm/{k}"].item()
        assert abs(gn - ref) <= grad_rel * 5 * ref + 1e-7, (k, gn, ref)
        a, b = sub(grads[k], 256), g[f"grad_sub/{k}"]
        assert cosine(a, b) > cos_min, (k, cosine(a, b))
        if post is not None:
            solid = np.abs(b) > 1e-2 * max(np.abs(b).max(), 1e-30)
            np.testing.assert_allclose(sub(post[k], 256)[solid], g[f"post_sub/{k}"][solid],
                                       atol=2e-4, rtol=1e-3, err_msg=k)
    if bufs is not None:
        for k, v in bufs.items():
            if "running_" in k:
                np.testing.assert_allclose(v.cpu().numpy(), g[f"post_buf/{k}"], rtol=2e-3, atol=2e-4,
                                           err_msg=k)


@pytest.mark.parametrize("name", list(CASES))
def test_train_step_autograd_fp32_matches_reference_golden(name):
    """Module path: loss.backward() + clip_grad_norm_ + torch.optim.Adam, dropout off, BN train."""
    from texttoaudiogrounding_b200.train import train_step
    g, sd, batch = load_case(name)
    model = build_model(sd, "fp32")
    model.train()
    model.audio_encoder.dropout_enabled = False
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    grads = {}

    def grab():
        for n, p in model.named_parameters():
            grads[n] = p.grad.detach().clone()
    # capture raw gradients before the in-place clip
    from texttoaudiogrounding_b200.train import runner_forward
    from texttoaudiogrounding_b200.losses import FrameBceLoss
    opt.zero_grad()
    out = runner_forward(model, dict(batch), "cuda", training=True)
    np.testing.assert_allclose(out["frame_sim"].detach().cpu().numpy(), g["train_frame_sim"], atol=1e-3)
    loss = FrameBceLoss()(out)
    loss.backward()
    grab()
    total_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    torch.cuda.synchronize()
    post = {n: p.detach() for n, p in model.named_parameters()}
    bufs = dict(model.named_buffers())
    _check_train_against_golden(g, grads, loss.item(), float(total_norm), post, bufs, 2e-3, 0.999)


@pytest.mark.parametrize("name", list(CASES))
def test_fused_train_step_fp32_matches_reference_golden(name):
    """Production path: flat buffers, fused clip+Adam kernel (eager and graph-replayed)."""
    from texttoaudiogrounding_b200.train import FusedTrainStep
    g, sd, batch = load_case(name)
    model = build_model(sd, "fp32")
    model.train()
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, lr=1e-3, max_grad_norm=1.0, use_graph=False)
    loss = ts.step(batch)
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    post = {n: p.detach() for n, p in model.named_parameters()}
    _check_train_against_golden(g, grads, loss.item(), ts.norm_out.item(), post,
                                dict(model.named_buffers()), 2e-3, 0.999)


def test_fused_graph_replay_equals_eager():
    from texttoaudiogrounding_b200.train import FusedTrainStep
    # un-sharpened weights: a well-conditioned problem, so that the fp32-atomic summation-order
    # noise of the split-K wgrad is not amplified by the optimisation dynamics
    sd = O.synth_state_dict(seed=1, sharpen=1.0, perturb_bn=True)
    batch = O.synth_batch(4, 64000, seed=0)
    losses = {}
    params = {}
    for use_graph in (False, True):
        model = build_model(sd, "fp32")
        model.train()
        model.audio_encoder.dropout_enabled = False
        ts = FusedTrainStep(model, use_graph=use_graph)
        ls = [ts.step(batch).item() for _ in range(6)]
        torch.cuda.synchronize()
        losses[use_graph] = ls
        params[use_graph] = ts.flat_p.clone()
        assert int(ts.step_dev.item()) == 6
    np.testing.assert_allclose(losses[True], losses[False], rtol=5e-3)
    assert losses[False][-1] < losses[False][0], losses[False]          # it learns the batch
    assert rel_err(params[True], params[False]) < 1e-2


def test_prefetched_inputs_and_side_stream_give_the_same_steps():
    """step(batch) after prefetch(batch) (copy stream + staging buffers, two alternating pinned host batches) and
    the side-stream weight-gradient schedule produce the same losses / parameters as the plain path."""
    from texttoaudiogrounding_b200.train import FusedTrainStep
    sd = O.synth_state_dict(seed=1, sharpen=1.0, perturb_bn=True)
    batches = [O.synth_batch(4, 64000, seed=s) for s in (0, 1)]
    pinned = [{k: (v.pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in b.items()} for b in batches]
    res = {}
    for mode in ("plain", "prefetch", "side"):
        model = build_model(sd, "fp32")
        model.train()
        model.audio_encoder.dropout_enabled = False
        ts = FusedTrainStep(model, use_graph=True)
        if mode == "side":
            ts.side_stream = torch.cuda.Stream()
        ls = []
        if mode == "prefetch":
            ts.prefetch(pinned[0])
        for i in range(6):
            b = pinned[i % 2] if mode == "prefetch" else batches[i % 2]
            loss = ts.step(b)
            if mode == "prefetch" and i + 1 < 6:
                ts.prefetch(pinned[(i + 1) % 2])
            ls.append(loss.item())
        torch.cuda.synchronize()
        res[mode] = (ls, ts.flat_p.clone())
    for mode in ("prefetch", "side"):
        np.testing.assert_allclose(res[mode][0], res["plain"][0], rtol=5e-3)
        assert rel_err(res[mode][1], res["plain"][1]) < 1e-2
    assert abs(res["plain"][0][0] - res["plain"][0][1]) > 1e-6          # the two batches differ


def test_train_step_bf16_close_to_reference():
    from texttoaudiogrounding_b200.train import FusedTrainStep
    g, sd, batch = load_case("cfg1_b4_2s")
    model = build_model(sd, "bf16")
    model.train()
    model.audio_encoder.dropout_enabled = False
    ts = FusedTrainStep(model, use_graph=False)
    loss = ts.step(batch)
    torch.cuda.synchronize()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters()}
    _check_train_against_golden(g, grads, loss.item(), ts.norm_out.item(), None, None, 3e-2, 0.95)


def test_dropout_training_forward_matches_oracle_with_same_masks():
    """Training-mode forward with dropout ON: export the kernel's masks and inject them in the oracle."""
    from texttoaudiogrounding_b200 import engine
    from texttoaudiogrounding_b200.ops import call
    g, sd, batch = load_case("cfg1_b4_2s")
    model = build_model(sd, "fp32")
    model.train()
    enc = model.audio_encoder
    seed = 1234
    Wt = enc._weights()
    wav = batch["waveform"].cuda()
    out, _ = engine.encoder_forward(Wt, wav, training=True, bn_training=True, dropout=True, seed=seed,
                                    dtype=torch.float32, save=False)
    B = wav.shape[0]
    shapes = {"block1": (B, 100, 32, 64), "block2": (B, 50, 16, 128), "block3": (B, 50, 8, 256),
              "block4": (B, 50, 4, 512)}
    masks = {}
    for i, (k, shp) in enumerate(shapes.items()):
        m = torch.empty(shp, device="cuda")
        call("tag_dropout_mask", m, m.numel(), 0.2, engine._seed_for(seed, i), None)
        masks[k] = nchw(m).cpu()
    m = torch.empty(B, 50, 512, device="cuda")
    call("tag_dropout_mask", m, m.numel(), 0.5, engine._seed_for(seed, 4), None)
    masks["fc_in"] = m.cpu()
    keep = float((masks["block1"] > 0).float().mean())
    assert abs(keep - 0.8) < 0.01, keep
    sd2 = {k: v.clone() for k, v in sd.items()}
    with torch.no_grad():
        ref = O.cnn8rnn_forward(sd2, batch["waveform"], batch["waveform_len"], training=True,
                                dropout_masks=masks)
    err = (out.cpu() - ref["embedding"]).abs().max().item()
    assert err < 2e-3, err


def test_graphs_are_cached_per_shape_and_learning_rate_is_live():
    """A collate that pads to the longest clip changes L between batches (datasets/collate_function.py:43-84):
    the graphs of every shape seen are kept (alternating shapes replay, no re-capture) and follow the eager
    trajectory; set_lr() acts on captured graphs (the kernel reads the device scalar)."""
    from texttoaudiogrounding_b200.train import FusedTrainStep
    sd = O.synth_state_dict(seed=1, sharpen=1.0, perturb_bn=True)
    batches = [O.synth_batch(4, 64000, seed=0), O.synth_batch(3, 48000, seed=1)]
    res = {}
    for use_graph in (False, True):
        model = build_model(sd, "fp32")
        model.train()
        model.audio_encoder.dropout_enabled = False
        ts = FusedTrainStep(model, use_graph=use_graph)
        ls = []
        for i in range(8):
            if i == 6:
                ts.set_lr(0.0)                      # from here on the parameters must not move
                frozen = ts.flat_p.clone()
            ls.append(ts.step(batches[i % 2]).item())
        torch.cuda.synchronize()
        assert torch.equal(ts.flat_p, frozen)
        assert int(ts.step_dev.item()) == 8
        if use_graph:
            assert len(ts._graphs) == 2             # one entry per shape, both still alive
        res[use_graph] = ls
    np.testing.assert_allclose(res[True], res[False], rtol=5e-3)
