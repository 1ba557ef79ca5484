"""Pin the CPU oracle (oracle/tag_oracle.py) against fixtures generated from the
unmodified reference (oracle/make_golden.py -> tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from oracle import tag_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = {
    "cfg1_b4_2s": dict(batch=4, n_samples=64000, ragged=False, seed=1, data_seed=0),
    "ragged_b4_1s": dict(batch=4, n_samples=32000, ragged=True, seed=2, data_seed=3),
}
SHARPEN = 300.0


def sub(t, n=512):
    flat = t.detach().reshape(-1)
    step = max(1, flat.numel() // n)
    return flat[::step][:n].float().numpy()


def load(name):
    cfg = CASES[name]
    g = np.load(os.path.join(GOLDEN, name + ".npz"))
    sd = O.synth_state_dict(seed=cfg["seed"], sharpen=SHARPEN, perturb_bn=True)
    batch = O.synth_batch(cfg["batch"], cfg["n_samples"], seed=cfg["data_seed"],
                          ragged=cfg["ragged"])
    # inputs re-synthesised from the seed must be the ones the fixture was made with
    assert np.array_equal(batch["text"].numpy(), g["text"])
    assert np.array_equal(batch["label"].numpy(), g["label"])
    assert np.array_equal(batch["waveform_len"], g["waveform_len"])
    np.testing.assert_allclose(batch["waveform"].double().sum().item(),
                               g["waveform_checksum"][0], rtol=1e-9, atol=1e-9)
    return g, sd, batch


def test_melscale_fbanks_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    ref = ta.functional.melscale_fbanks(513, 50.0, 14000.0, 64, 32000, norm="slaney",
                                        mel_scale="slaney")
    np.testing.assert_allclose(O.melscale_fbanks().numpy(), ref.numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(O.hann_window().numpy(), torch.hann_window(1024).numpy(),
                               atol=5e-7)


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_eval_forward_matches_reference(name):
    g, sd, batch = load(name)
    stages = {}
    with torch.no_grad():
        out = O.runner_forward(sd, batch, training=False, stages=stages)
    db = stages["logmel_db"].numpy()
    # |dB| ~ 10..100; fp32 FFT summation-order noise only
    np.testing.assert_allclose(db, g["eval_logmel_db"], rtol=0, atol=2e-3)
    for nm in ["bn0", "conv_block1", "conv_block2", "conv_block3", "conv_block4", "fc1"]:
        np.testing.assert_allclose(sub(stages[nm]), g[f"eval_{nm}_sub"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(stages["rnn"].numpy(), g["eval_embedding"], atol=1e-3)
    np.testing.assert_allclose(stages["logits"].numpy(), g["eval_logits"], atol=2e-3, rtol=1e-3)
    # contractual bar: <= 1e-3 on the frame-probability tensor (fp32)
    assert np.abs(out["frame_sim"].numpy() - g["eval_frame_sim"]).max() <= 1e-3
    assert np.array_equal(out["length"].numpy(), g["eval_length"])


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_train_step_matches_reference(name):
    g, sd, batch = load(name)
    keys = O.trainable_keys()
    assert sorted(keys) == sorted(g["param_names"].tolist())
    opt = O.AdamState(keys)
    loss, grads, total_norm = O.train_step(sd, batch, opt, lr=1e-3, max_grad_norm=1.0,
                                           dropout=False)
    np.testing.assert_allclose(loss.item(), g["train_loss"].item(), rtol=1e-4)
    np.testing.assert_allclose(total_norm.item(), g["train_total_norm"].item(), rtol=2e-3)
    for k in keys:
        gn = grads[k].double().pow(2).sum().sqrt().item()
        ref = g[f"grad_norm/{k}"].item()
        assert abs(gn - ref) <= 5e-3 * ref + 1e-7, (k, gn, ref)
        a, b = sub(grads[k], 256), g[f"grad_sub/{k}"]
        denom = np.linalg.norm(a) * np.linalg.norm(b)
        if denom > 0:
            assert (a * b).sum() / denom > 0.999, k
        # Adam's first step is lr*sign(g) (eps 1e-8): elements whose gradient is at the
        # fp32 noise floor can legitimately land on either side, so compare the rest.
        solid = np.abs(b) > 1e-2 * max(np.abs(b).max(), 1e-30)
        np.testing.assert_allclose(sub(sd[k], 256)[solid], g[f"post_sub/{k}"][solid],
                                   atol=2e-4, rtol=1e-3)
    for k in sd:
        if "running_" in k:
            np.testing.assert_allclose(sd[k].numpy(), g[f"post_buf/{k}"], rtol=1e-3, atol=1e-4)


def test_fast_gru_equals_loop():
    sd = O.synth_state_dict(seed=5)
    x = torch.randn(3, 17, 512, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        a = O.bigru(x, sd, fast=False)
        b = O.bigru(x, sd, fast=True)
    np.testing.assert_allclose(a.numpy(), b.numpy(), atol=2e-6)


def test_oracle_matches_reference_at_full_clip_length():
    """The 10 s / ragged full-size fixture (oracle/make_golden_autocast.py ran the unmodified reference on all 64 clips):
    eval is per-clip independent, so four of its rows pin the oracle at the BASELINE clip length within seconds."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "autocast_b64_10s.npz"))
    batch = O.synth_batch(64, 320000, seed=21, ragged=True)
    rows = [0, 1, 2, 63]
    sub_batch = {k: (v[rows] if not isinstance(v, list) else [v[i] for i in rows]) for k, v in batch.items()}
    sd = O.synth_state_dict(seed=3, sharpen=300.0, perturb_bn=True)
    with torch.no_grad():
        out = O.runner_forward(sd, sub_batch, training=False, fast_gru=True)["frame_sim"].numpy()
    assert np.abs(out - g["eval_frame_sim_fp32"][rows]).max() <= 1e-4
    # what the reference itself loses under torch.autocast(bfloat16) on these inputs: the yardstick of the bf16 tests
    assert np.abs(g["eval_frame_sim_autocast"] - g["eval_frame_sim_fp32"]).max() > 5e-2
    assert float(g["grad_cosine/audio_encoder.bn0.weight"]) < 0.96
